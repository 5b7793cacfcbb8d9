// Decoder stage entry points: Phi3ForCausalLM / LlamaForCausalLM prefill + greedy decode
// (modeling_phi3.py:1249-1383 Phi3Model.forward, :1034-1095 Phi3DecoderLayer, :629-775 attention,
//  :458-464 MLP, :1525-1526 lm_head + .float(); modeling_llama.py:934-1044, 699-760, 417-594, 218-238;
//  the HF GenerationMixin greedy loop invoked at llava_next_video.py:655-661).
//
// The object owns a pre-allocated KV cache [layer][k|v][kv_heads][max_ctx][head_dim] (the reference
// grows a DynamicCache with torch.cat every step, modeling_phi3.py:721), the activation workspace,
// and a CUDA graph of one decode step. All step bookkeeping (position, current token, EOS flag)
// lives in device memory so the graph is replayed without host round trips.
#include <vector>
#include "gvl_internal.h"
#include <stdlib.h>
#include "decode.h"
#include "decode_mega.h"
#include "../../include/gvl.h"

using namespace gvl;

struct gvl_lm {
    gvl_lm_weights w;
    std::vector<gvl_lm_layer> layers;
    __nv_bfloat16* kv = nullptr;      // [L][2][KVH][max_ctx][hd]
    // prefill workspace (grown on demand)
    __nv_bfloat16 *x = nullptr, *h = nullptr, *qkv = nullptr, *q = nullptr, *attn = nullptr, *mid = nullptr;
    int ws_tokens = 0;
    // decode workspace
    __nv_bfloat16 *dx = nullptr, *dqkv = nullptr, *dq = nullptr, *dattn = nullptr, *dmid = nullptr;
    float* dlogits = nullptr;
    float* da_ws = nullptr;
    DecodeState* st = nullptr;
    long long* first_tok = nullptr;
    // decode graph (captured per (tokens_out, logits_out, eos, pad) binding)
    cudaGraphExec_t graph = nullptr;
    float* g_logits = nullptr;
    long long* tok_buf = nullptr;     // [max_ctx] generated tokens of the current decode call
    float* logit_buf = nullptr;       // [logit_steps][vocab], only when the caller asks for logits
    int logit_steps = 0;
    cudaStream_t cs = nullptr;        // capture / replay stream
    cudaEvent_t ev_in = nullptr, ev_out = nullptr;
    bool warmed = false;
    long long g_eos = 0, g_pad = 0;
    bool use_graph = true;
    bool use_pdl = true;
    int host_ctx = -1;                // tokens in the KV cache as the HOST knows them: S after prefill, += n_steps per decode call
    // single-kernel decode step (decode_mega.cu): the default when the shape fits (GVL_DECODE_MEGA=0 selects the per-op chain)
    bool use_mega = false;
    MegaPlan* plan_dev = nullptr;
    MegaPlan* plan_host = nullptr;
    unsigned* grid_bar = nullptr;
    float* mega_att_ws = nullptr;
    __nv_bfloat16* mega_w = nullptr;  // packed decode copy of every layer's weights + lm_head (decode_mega_pack)
    unsigned long long* amax = nullptr;
    long long* trace = nullptr;       // GVL_MEGA_TRACE=1: per-CTA phase timestamps of the last decode step
    // batched decode (gvl_lm_decode_batch): activation rows of up to GVL_LM_MAX_BATCH sequences side by side, owned by the FIRST object
    // of a batch; one CUDA graph per batch composition
    __nv_bfloat16 *bx = nullptr, *bqkv = nullptr, *bq = nullptr, *battn = nullptr, *bmid = nullptr;
    float* blogits = nullptr;
    cudaGraphExec_t bgraph = nullptr;
    gvl_lm* b_members[4] = {nullptr, nullptr, nullptr, nullptr};
    int b_n = 0;
    bool b_logits = false;
    long long b_eos = 0, b_pad = 0;
    unsigned b_warm_mask = 0;         // batch sizes whose kernels have run once outside capture (function attributes set)

    size_t kv_layer_elems() const { return (size_t)2 * w.kv_heads * w.max_ctx * w.head_dim; }
    __nv_bfloat16* kcache(int l) const { return kv + (size_t)l * kv_layer_elems(); }
    __nv_bfloat16* vcache(int l) const { return kcache(l) + (size_t)w.kv_heads * w.max_ctx * w.head_dim; }
};

namespace {

#define CK(expr)                       \
    do {                               \
        int _rc = (expr);              \
        if (_rc != GVL_OK) return _rc; \
    } while (0)
#define CU(expr)                                   \
    do {                                           \
        if ((expr) != cudaSuccess) return GVL_ERR_CUDA; \
    } while (0)

template <typename T>
int dev_alloc(T** p, size_t n) {
    return cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(T)) == cudaSuccess ? GVL_OK : GVL_ERR_NOMEM;
}

int ensure_prefill_ws(gvl_lm* lm, int S) {
    if (S <= lm->ws_tokens) return GVL_OK;
    const gvl_lm_weights& w = lm->w;
    cudaFree(lm->x); cudaFree(lm->h); cudaFree(lm->qkv); cudaFree(lm->q); cudaFree(lm->attn); cudaFree(lm->mid);
    lm->x = lm->h = lm->qkv = lm->q = lm->attn = lm->mid = nullptr;
    lm->ws_tokens = 0;
    const size_t qkv_n = (size_t)(w.heads + 2 * w.kv_heads) * w.head_dim;
    CK(dev_alloc(&lm->x, (size_t)S * w.dim));
    CK(dev_alloc(&lm->h, (size_t)S * w.dim));
    CK(dev_alloc(&lm->qkv, (size_t)S * qkv_n));
    CK(dev_alloc(&lm->q, (size_t)S * w.heads * w.head_dim));
    CK(dev_alloc(&lm->attn, (size_t)S * w.heads * w.head_dim));
    CK(dev_alloc(&lm->mid, (size_t)S * w.ffn));
    lm->ws_tokens = S;
    return GVL_OK;
}

// One decode step on `s` (graph-capturable: no host-dependent arguments change between steps).
int enqueue_decode_step_impl(gvl_lm* lm, long long* tokens_out, float* logits_out, long long eos_id,
                             long long pad_id, cudaStream_t s) {
    const gvl_lm_weights& w = lm->w;
    const int D = w.dim, H = w.heads, KVH = w.kv_heads, hd = w.head_dim, F = w.ffn;
    const int qkv_n = (H + 2 * KVH) * hd;
    const float scale = 1.0f / sqrtf((float)hd);
    CK(step_begin(lm->st, s));
    CK(embed_token((const __nv_bfloat16*)w.embed, lm->st, lm->dx, D, w.vocab, s));
    for (int l = 0; l < w.n_layers; ++l) {
        const gvl_lm_layer& L = lm->layers[l];
        CK(gemv_bf16(lm->dx, D, (const __nv_bfloat16*)L.qkv_w, D, lm->dqkv, qkv_n, 1, qkv_n, D,
                     (const __nv_bfloat16*)L.in_norm_w, w.rms_eps, nullptr, nullptr, 0, 0, 0, s));
        CK(rope_decode(lm->dqkv, lm->dq, lm->kcache(l), lm->vcache(l), (const __nv_bfloat16*)w.rope_cos,
                       (const __nv_bfloat16*)w.rope_sin, lm->st, H, KVH, hd, w.max_ctx, s));
        CK(decode_attention(lm->dq, lm->kcache(l), lm->vcache(l), lm->dattn, lm->da_ws, &lm->st->attn_len, H, KVH, hd,
                            w.max_ctx, scale, s));
        CK(gemv_bf16(lm->dattn, H * hd, (const __nv_bfloat16*)L.o_w, H * hd, lm->dx, D, 1, D, H * hd, nullptr, 0.f,
                     nullptr, lm->dx, D, 0, 0, s));
        CK(gemv_bf16(lm->dx, D, (const __nv_bfloat16*)L.gate_up_w, D, lm->dmid, F, 1, 2 * F, D,
                     (const __nv_bfloat16*)L.post_norm_w, w.rms_eps, nullptr, nullptr, 0, 3, 0, s));
        CK(gemv_bf16(lm->dmid, F, (const __nv_bfloat16*)L.down_w, F, lm->dx, D, 1, D, F, nullptr, 0.f, nullptr,
                     lm->dx, D, 0, 0, s));
    }
    CK(gemv_bf16(lm->dx, D, (const __nv_bfloat16*)w.lm_head_w, D, lm->dlogits, w.vocab, 1, w.vocab, D,
                 (const __nv_bfloat16*)w.final_norm_w, w.rms_eps, (const __nv_bfloat16*)w.lm_head_b, nullptr, 0, 0, 1,
                 s));
    CK(step_end(lm->dlogits, w.vocab, lm->st, tokens_out, logits_out, eos_id, pad_id, s));
    return GVL_OK;
}

// PDL is enabled for the decode chain only (every kernel in it starts with pdl_wait()).
int enqueue_decode_step(gvl_lm* lm, long long* tokens_out, float* logits_out, long long eos_id, long long pad_id,
                        cudaStream_t s) {
    if (lm->use_mega && lm->plan_dev)
        return decode_mega_launch(lm->plan_host, lm->plan_dev, 1, tokens_out, logits_out, eos_id, pad_id, s);
    g_pdl = lm->use_pdl;
    const int rc = enqueue_decode_step_impl(lm, tokens_out, logits_out, eos_id, pad_id, s);
    g_pdl = false;
    return rc;
}

}  // namespace

extern "C" {

int gvl_lm_create(const gvl_lm_weights* w, gvl_lm** out) { return gvl_lm_create_ex(w, 0, out); }

int gvl_lm_create_ex(const gvl_lm_weights* w, int flags, gvl_lm** out) {
    if (!w || !out || w->n_layers <= 0 || !w->layers) return GVL_ERR_ARG;
    if (w->dim % 256 != 0 || w->ffn % 256 != 0 || (w->heads * w->head_dim) % 256 != 0) return GVL_ERR_ARG;
    if (w->head_dim != 64 && w->head_dim != 96 && w->head_dim != 128) return GVL_ERR_ARG;
    gvl_lm* lm = new gvl_lm();
    lm->w = *w;
    lm->layers.assign(w->layers, w->layers + w->n_layers);
    lm->w.layers = lm->layers.data();
    const int D = w->dim, qkv_n = (w->heads + 2 * w->kv_heads) * w->head_dim;
    int rc = GVL_OK;
    if ((rc = dev_alloc(&lm->kv, (size_t)w->n_layers * lm->kv_layer_elems())) != GVL_OK) goto fail;
    if ((rc = dev_alloc(&lm->dx, (size_t)D)) != GVL_OK) goto fail;
    if ((rc = dev_alloc(&lm->dqkv, (size_t)qkv_n)) != GVL_OK) goto fail;
    if ((rc = dev_alloc(&lm->dq, (size_t)w->heads * w->head_dim)) != GVL_OK) goto fail;
    if ((rc = dev_alloc(&lm->dattn, (size_t)w->heads * w->head_dim)) != GVL_OK) goto fail;
    if ((rc = dev_alloc(&lm->dmid, (size_t)w->ffn)) != GVL_OK) goto fail;
    if ((rc = dev_alloc(&lm->dlogits, (size_t)w->vocab)) != GVL_OK) goto fail;
    if ((rc = dev_alloc(&lm->da_ws, decode_attention_workspace(w->heads, w->head_dim, w->max_ctx) / sizeof(float))) != GVL_OK) goto fail;
    if ((rc = dev_alloc(&lm->st, 1)) != GVL_OK) goto fail;
    if ((rc = dev_alloc(&lm->first_tok, 1)) != GVL_OK) goto fail;
    if ((rc = dev_alloc(&lm->tok_buf, (size_t)w->max_ctx)) != GVL_OK) goto fail;
    if (cudaStreamCreateWithFlags(&lm->cs, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&lm->ev_in, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&lm->ev_out, cudaEventDisableTiming) != cudaSuccess) { rc = GVL_ERR_CUDA; goto fail; }
    cudaMemset(lm->st, 0, sizeof(DecodeState));
    cudaMemset(lm->da_ws, 0, decode_attention_workspace(w->heads, w->head_dim, w->max_ctx));
    {
        const char* env = getenv("GVL_DECODE_MEGA");
        // single-kernel decode step is the default (measured faster than the per-op chain); GVL_DECODE_MEGA=0 selects the chain
        lm->use_mega = !(env && env[0] == '0') && !(flags & GVL_LM_NO_SINGLE_KERNEL) && w->n_layers <= MEGA_MAX_LAYERS;
        if (lm->use_mega) {
            MegaPlan* hp = new MegaPlan();
            memset(hp, 0, sizeof(MegaPlan));
            hp->n_layers = w->n_layers; hp->dim = w->dim; hp->heads = w->heads; hp->kv_heads = w->kv_heads;
            hp->head_dim = w->head_dim; hp->vocab = w->vocab; hp->max_ctx = w->max_ctx;
            hp->scale = 1.0f / sqrtf((float)w->head_dim);
            bool ok = true;
            auto mk = [&](const void* W, int N, int K, const __nv_bfloat16* x, const void* norm_w, const void* bias,
                          const __nv_bfloat16* residual, void* out, int act, int out_f32) {
                MegaOp op;
                memset(&op, 0, sizeof(op));
                op.W = (const __nv_bfloat16*)W; op.K = K; op.n_rows = N;      // W: source for now, packed below
                op.act = act; op.out_f32 = out_f32; op.x = x; op.norm_w = (const __nv_bfloat16*)norm_w; op.eps = w->rms_eps;
                op.bias = (const __nv_bfloat16*)bias; op.residual = residual; op.out = out;
                ok = decode_mega_shape(&op) && ok;
                return op;
            };
            const int HD = w->heads * w->head_dim;
            for (int l = 0; l < w->n_layers; ++l) {
                const gvl_lm_layer& L = lm->layers[l];
                hp->ops[l * 4 + 0] = mk(L.qkv_w, qkv_n, D, lm->dx, L.in_norm_w, nullptr, nullptr, lm->dqkv, 0, 0);
                hp->ops[l * 4 + 1] = mk(L.o_w, D, HD, nullptr, nullptr, nullptr, lm->dx, lm->dx, 0, 0);
                hp->ops[l * 4 + 1].x_kind = 1;
                hp->ops[l * 4 + 2] = mk(L.gate_up_w, 2 * w->ffn, D, lm->dx, L.post_norm_w, nullptr, nullptr, lm->dmid, 3, 0);
                hp->ops[l * 4 + 3] = mk(L.down_w, D, w->ffn, lm->dmid, nullptr, nullptr, lm->dx, lm->dx, 0, 0);
            }
            hp->ops[0].from_embed = 1;          // layer 0 reads the token's embedding row directly (no embed phase)
            hp->ops[1].from_embed = 2;
            hp->ops[w->n_layers * 4] = mk(w->lm_head_w, w->vocab, D, lm->dx, w->final_norm_w, w->lm_head_b, nullptr, lm->dlogits, 0, 1);
            hp->ops[w->n_layers * 4].argmax = 1;
            hp->embed = (const __nv_bfloat16*)w->embed;
            hp->rope_cos = (const __nv_bfloat16*)w->rope_cos; hp->rope_sin = (const __nv_bfloat16*)w->rope_sin;
            hp->kv = lm->kv; hp->x = lm->dx; hp->qkv = lm->dqkv; hp->mid = lm->dmid;
            hp->logits = lm->dlogits;
            hp->st = lm->st;
            ok = ok && decode_mega_finalize(hp);
            if (!ok) {                           // shape not covered by the single-kernel step: per-op chain
                delete hp;
                lm->use_mega = false;
            } else {
                hp->trace = nullptr;
                if (getenv("GVL_MEGA_TRACE") && dev_alloc(&lm->trace, (size_t)num_sms() * MEGA_TRACE_STRIDE) == GVL_OK) {
                    cudaMemset(lm->trace, 0, (size_t)num_sms() * MEGA_TRACE_STRIDE * sizeof(long long));
                    hp->trace = lm->trace;
                }
                const int n_ops = w->n_layers * 4 + 1;
                size_t pack_total = 0;
                for (int i = 0; i < n_ops; ++i) pack_total += decode_mega_packed_elems(&hp->ops[i]);
                bool aok = dev_alloc(&lm->mega_w, pack_total) == GVL_OK;
                if (aok) {
                    size_t off = 0;
                    for (int i = 0; i < n_ops && aok; ++i) {
                        MegaOp& op = hp->ops[i];
                        aok = decode_mega_pack(&op, op.W, op.K, lm->mega_w + off, 0) == GVL_OK;
                        op.W = lm->mega_w + off;
                        off += decode_mega_packed_elems(&op);
                    }
                    aok = aok && cudaDeviceSynchronize() == cudaSuccess;
                }
                aok = aok && dev_alloc(&lm->grid_bar, 1) == GVL_OK && dev_alloc(&lm->plan_dev, 1) == GVL_OK &&
                           dev_alloc(&lm->amax, 1) == GVL_OK &&
                           dev_alloc(&lm->mega_att_ws, decode_mega_att_ws_bytes(hp) / sizeof(float)) == GVL_OK;
                if (aok) {
                    hp->grid_bar = lm->grid_bar;
                    hp->amax = lm->amax;
                    hp->att_ws = lm->mega_att_ws;
                    aok = cudaMemset(lm->amax, 0, sizeof(unsigned long long)) == cudaSuccess &&
                          cudaMemcpy(lm->plan_dev, hp, sizeof(MegaPlan), cudaMemcpyHostToDevice) == cudaSuccess;
                }
                lm->plan_host = hp;
                if (!aok) { rc = GVL_ERR_NOMEM; goto fail; }
            }
        }
    }
    *out = lm;
    return GVL_OK;
fail:
    gvl_lm_destroy(lm);
    return rc;
}

void gvl_lm_destroy(gvl_lm* lm) {
    if (!lm) return;
    if (lm->graph) cudaGraphExecDestroy(lm->graph);
    cudaFree(lm->kv);
    cudaFree(lm->x); cudaFree(lm->h); cudaFree(lm->qkv); cudaFree(lm->q); cudaFree(lm->attn); cudaFree(lm->mid);
    cudaFree(lm->dx); cudaFree(lm->dqkv); cudaFree(lm->dq); cudaFree(lm->dattn); cudaFree(lm->dmid);
    cudaFree(lm->dlogits); cudaFree(lm->da_ws); cudaFree(lm->st); cudaFree(lm->first_tok);
    cudaFree(lm->tok_buf); cudaFree(lm->logit_buf); cudaFree(lm->plan_dev); cudaFree(lm->grid_bar); cudaFree(lm->trace);
    cudaFree(lm->mega_att_ws); cudaFree(lm->amax); cudaFree(lm->mega_w);
    if (lm->bgraph) cudaGraphExecDestroy(lm->bgraph);
    cudaFree(lm->bx); cudaFree(lm->bqkv); cudaFree(lm->bq); cudaFree(lm->battn); cudaFree(lm->bmid); cudaFree(lm->blogits);
    delete lm->plan_host;
    if (lm->cs) cudaStreamDestroy(lm->cs);
    if (lm->ev_in) cudaEventDestroy(lm->ev_in);
    if (lm->ev_out) cudaEventDestroy(lm->ev_out);
    delete lm;
}

int gvl_lm_prefill(gvl_lm* lm, const void* embeds, int S, float* logits_out, void* hidden_out, void* stream) {
    if (!lm || !embeds || S <= 0 || S > lm->w.max_ctx) return GVL_ERR_ARG;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const gvl_lm_weights& w = lm->w;
    CK(ensure_prefill_ws(lm, S));
    const int D = w.dim, H = w.heads, KVH = w.kv_heads, hd = w.head_dim, F = w.ffn;
    const int qkv_n = (H + 2 * KVH) * hd;
    const float scale = 1.0f / sqrtf((float)hd);
    CU(cudaMemcpyAsync(lm->x, embeds, (size_t)S * D * 2, cudaMemcpyDeviceToDevice, s));
    for (int l = 0; l < w.n_layers; ++l) {
        const gvl_lm_layer& L = lm->layers[l];
        // Phi3DecoderLayer.forward (modeling_phi3.py:1034-1095)
        CK(rmsnorm_bf16(lm->x, D, (const __nv_bfloat16*)L.in_norm_w, lm->h, D, S, D, w.rms_eps, s));
        CK(gemm_bf16(lm->h, D, L.qkv_w, D, lm->qkv, qkv_n, S, qkv_n, D, nullptr, nullptr, nullptr, 0, GVL_ACT_NONE,
                     GVL_RES_NONE, 0, 0, s));
        CK(rope_qkv_cache(lm->qkv, lm->q, lm->kcache(l), lm->vcache(l), (const __nv_bfloat16*)w.rope_cos,
                          (const __nv_bfloat16*)w.rope_sin, nullptr, S, H, KVH, hd, 0, w.max_ctx, s));
        AttnArgs a;
        a.q = lm->q; a.k = lm->kcache(l); a.v = lm->vcache(l); a.o = lm->attn;
        a.q_bs = 0; a.q_ts = (long long)H * hd; a.q_hs = hd;
        a.k_bs = a.v_bs = 0; a.k_ts = a.v_ts = hd; a.k_hs = a.v_hs = (long long)w.max_ctx * hd;
        a.o_bs = 0; a.o_ts = (long long)H * hd; a.o_hs = hd;
        a.batch = 1; a.heads = H; a.kv_heads = KVH; a.sq = S; a.skv = S; a.head_dim = hd;
        a.scale = scale; a.causal = 1; a.round_scores = 0;
        CK(attention_fwd(a, s));
        CK(gemm_bf16(lm->attn, H * hd, L.o_w, H * hd, lm->x, D, S, D, H * hd, nullptr, nullptr, lm->x, D, GVL_ACT_NONE,
                     GVL_RES_BF16, 0, 0, s));
        CK(rmsnorm_bf16(lm->x, D, (const __nv_bfloat16*)L.post_norm_w, lm->h, D, S, D, w.rms_eps, s));
        CK(gemm_bf16(lm->h, D, L.gate_up_w, D, lm->mid, F, S, 2 * F, D, nullptr, nullptr, nullptr, 0, GVL_ACT_SWIGLU,
                     GVL_RES_NONE, 0, 0, s));
        CK(gemm_bf16(lm->mid, F, L.down_w, F, lm->x, D, S, D, F, nullptr, nullptr, lm->x, D, GVL_ACT_NONE, GVL_RES_BF16,
                     0, 0, s));
    }
    if (hidden_out) CU(cudaMemcpyAsync(hidden_out, lm->x, (size_t)S * D * 2, cudaMemcpyDeviceToDevice, s));
    // final norm + lm_head on the LAST position only (the reference computes all S rows and reads one,
    // modeling_phi3.py:1525-1526)
    CK(gemv_bf16(lm->x + (size_t)(S - 1) * D, D, (const __nv_bfloat16*)w.lm_head_w, D, lm->dlogits, w.vocab, 1, w.vocab,
                 D, (const __nv_bfloat16*)w.final_norm_w, w.rms_eps, (const __nv_bfloat16*)w.lm_head_b, nullptr, 0, 0, 1,
                 s));
    if (logits_out) CU(cudaMemcpyAsync(logits_out, lm->dlogits, (size_t)w.vocab * 4, cudaMemcpyDeviceToDevice, s));
    CK(argmax_f32(lm->dlogits, w.vocab, lm->first_tok, s));
    DecodeState init;
    init.ctx_len = S; init.attn_len = S; init.step = 0; init.finished = 0; init.cur_token = 0;
    CU(cudaMemcpyAsync(lm->st, &init, sizeof(init), cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(&lm->st->cur_token, lm->first_tok, sizeof(long long), cudaMemcpyDeviceToDevice, s));
    lm->host_ctx = S;
    return GVL_OK;
}

int gvl_lm_decode(gvl_lm* lm, int n_steps, long long* tokens_out, float* logits_out, long long eos_id,
                  long long pad_id, void* stream) {
    if (!lm || n_steps < 0 || n_steps > lm->w.max_ctx) return GVL_ERR_ARG;
    cudaStream_t caller = reinterpret_cast<cudaStream_t>(stream);
    if (lm->host_ctx < 0) return GVL_ERR_STATE;                         // no prefill yet
    if (lm->host_ctx + n_steps > lm->w.max_ctx) return GVL_ERR_STATE;   // step t writes K/V slot ctx + t: never past the cache
    if (n_steps == 0) return GVL_OK;
    // Steps run on the object's own stream (the caller's may be the legacy default stream, which cannot be
    // captured), fenced against the caller's stream with events on both sides.
    cudaStream_t s = lm->cs;
    CU(cudaEventRecord(lm->ev_in, caller));
    CU(cudaStreamWaitEvent(s, lm->ev_in, 0));
    // results land in internal buffers so that the captured graph does not depend on caller pointers
    const bool want_logits = logits_out != nullptr;
    if (want_logits && lm->logit_steps < n_steps) {
        cudaFree(lm->logit_buf);
        lm->logit_buf = nullptr;
        CK(dev_alloc(&lm->logit_buf, (size_t)n_steps * lm->w.vocab));
        lm->logit_steps = n_steps;
        if (lm->graph) { cudaGraphExecDestroy(lm->graph); lm->graph = nullptr; }
    }
    float* lbuf = want_logits ? lm->logit_buf : nullptr;
    int zero = 0;
    CU(cudaMemcpyAsync(&lm->st->step, &zero, sizeof(int), cudaMemcpyHostToDevice, s));
    if (lm->use_mega && lm->plan_dev) {
        // one cooperative launch runs all the steps (token / position / EOS state stay in device memory)
        CK(decode_mega_launch(lm->plan_host, lm->plan_dev, n_steps, lm->tok_buf, lbuf, eos_id, pad_id, s));
    } else if (!lm->use_graph) {
        for (int i = 0; i < n_steps; ++i) CK(enqueue_decode_step(lm, lm->tok_buf, lbuf, eos_id, pad_id, s));
    } else {
        if (lm->graph == nullptr || lm->g_logits != lbuf || lm->g_eos != eos_id || lm->g_pad != pad_id) {
            if (lm->graph) { cudaGraphExecDestroy(lm->graph); lm->graph = nullptr; }
            if (!lm->warmed) {
                // one eager step outside capture so that cudaFuncSetAttribute is never issued while capturing;
                // its state mutation is undone by restoring the DecodeState (the KV slot is rewritten later).
                DecodeState saved;
                CU(cudaMemcpyAsync(&saved, lm->st, sizeof(saved), cudaMemcpyDeviceToHost, s));
                CU(cudaStreamSynchronize(s));
                CK(enqueue_decode_step(lm, nullptr, nullptr, -1, 0, s));
                CU(cudaMemcpyAsync(lm->st, &saved, sizeof(saved), cudaMemcpyHostToDevice, s));
                CU(cudaStreamSynchronize(s));
                lm->warmed = true;
            }
            cudaGraph_t g = nullptr;
            CU(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
            const long long launches_before = g_launch_count;
            int rc = enqueue_decode_step(lm, lm->tok_buf, lbuf, eos_id, pad_id, s);
            cudaError_t e = cudaStreamEndCapture(s, &g);
            g_launch_count = launches_before;  // captured launches did not execute
            if (rc != GVL_OK || e != cudaSuccess) { if (g) cudaGraphDestroy(g); return rc != GVL_OK ? rc : GVL_ERR_CUDA; }
            e = cudaGraphInstantiate(&lm->graph, g, 0);
            cudaGraphDestroy(g);
            if (e != cudaSuccess) return GVL_ERR_CUDA;
            lm->g_logits = lbuf; lm->g_eos = eos_id; lm->g_pad = pad_id;
        }
        const long long per_step = (lm->use_mega && lm->plan_dev) ? 1 : 2 + (long long)lm->w.n_layers * 6 + 2;
        for (int i = 0; i < n_steps; ++i) {
            CU(cudaGraphLaunch(lm->graph, s));
            g_launch_count += per_step;
        }
    }
    if (tokens_out) CU(cudaMemcpyAsync(tokens_out, lm->tok_buf, (size_t)n_steps * sizeof(long long), cudaMemcpyDeviceToDevice, s));
    if (want_logits) CU(cudaMemcpyAsync(logits_out, lbuf, (size_t)n_steps * lm->w.vocab * sizeof(float), cudaMemcpyDeviceToDevice, s));
    CU(cudaEventRecord(lm->ev_out, s));
    CU(cudaStreamWaitEvent(caller, lm->ev_out, 0));
    lm->host_ctx += n_steps;
    return GVL_OK;
}

// ---------------------------------------------------------------------------------------------------------------------------
// Batched greedy decode: n_seq <= 4 sequences that were prefilled on their OWN gvl_lm objects (own KV cache, own decode state, own RoPE
// table) over the SAME weights advance together. Every weight matrix is streamed once per step for all sequences (gemv3_kernel with
// M = n_seq activation rows: the weights are 7.4 of the 8.8 GB a Phi-3.5 step moves); RoPE + KV append, the q_len = 1 attention and the
// greedy pick stay per sequence. Semantics per sequence are those of gvl_lm_decode (HF greedy: EOS -> pad afterwards).
// Reference: the batched `language_model.generate` of llava_next_video.py:655-661 (rows are independent; positions per row).
static int enqueue_batch_step(gvl_lm* const* lms, int n, long long eos_id, long long pad_id, bool want_logits, cudaStream_t s) {
    gvl_lm* L0 = lms[0];
    const gvl_lm_weights& w = L0->w;
    const int D = w.dim, H = w.heads, KVH = w.kv_heads, hd = w.head_dim, F = w.ffn;
    const int HD = H * hd, qkv_n = (H + 2 * KVH) * hd;
    const float scale = 1.0f / sqrtf((float)hd);
    for (int b = 0; b < n; ++b) {
        CK(step_begin(lms[b]->st, s));
        CK(embed_token((const __nv_bfloat16*)w.embed, lms[b]->st, L0->bx + (size_t)b * D, D, w.vocab, s));
    }
    for (int l = 0; l < w.n_layers; ++l) {
        const gvl_lm_layer& L = L0->layers[l];
        CK(gemv_bf16(L0->bx, D, (const __nv_bfloat16*)L.qkv_w, D, L0->bqkv, qkv_n, n, qkv_n, D, (const __nv_bfloat16*)L.in_norm_w,
                     w.rms_eps, nullptr, nullptr, 0, 0, 0, s));
        // RoPE + append and the split-KV attention of ALL sequences in one launch each (own cache / position / workspace per sequence)
        DecodeRopeBatch rb = {};
        DecodeAttnBatch ab = {};
        for (int b = 0; b < n; ++b) {
            gvl_lm* m = lms[b];
            rb.s[b] = {L0->bqkv + (size_t)b * qkv_n, L0->bq + (size_t)b * HD, m->kcache(l), m->vcache(l),
                       (const __nv_bfloat16*)m->w.rope_cos, (const __nv_bfloat16*)m->w.rope_sin, m->st, m->w.max_ctx};
            ab.s[b] = {L0->bq + (size_t)b * HD, m->kcache(l), m->vcache(l), L0->battn + (size_t)b * HD, m->da_ws, &m->st->attn_len,
                       m->w.max_ctx};
        }
        CK(rope_decode_batch(rb, n, H, KVH, hd, s));
        CK(decode_attention_batch(ab, n, H, KVH, hd, scale, s));
        CK(gemv_bf16(L0->battn, HD, (const __nv_bfloat16*)L.o_w, HD, L0->bx, D, n, D, HD, nullptr, 0.f, nullptr, L0->bx, D, 0, 0, s));
        CK(gemv_bf16(L0->bx, D, (const __nv_bfloat16*)L.gate_up_w, D, L0->bmid, F, n, 2 * F, D, (const __nv_bfloat16*)L.post_norm_w,
                     w.rms_eps, nullptr, nullptr, 0, 3, 0, s));
        CK(gemv_bf16(L0->bmid, F, (const __nv_bfloat16*)L.down_w, F, L0->bx, D, n, D, F, nullptr, 0.f, nullptr, L0->bx, D, 0, 0, s));
    }
    CK(gemv_bf16(L0->bx, D, (const __nv_bfloat16*)w.lm_head_w, D, L0->blogits, w.vocab, n, w.vocab, D, (const __nv_bfloat16*)w.final_norm_w,
                 w.rms_eps, (const __nv_bfloat16*)w.lm_head_b, nullptr, 0, 0, 1, s));
    for (int b = 0; b < n; ++b)
        CK(step_end(L0->blogits + (size_t)b * w.vocab, w.vocab, lms[b]->st, lms[b]->tok_buf, want_logits ? lms[b]->logit_buf : nullptr,
                    eos_id, pad_id, s));
    return GVL_OK;
}

int gvl_lm_decode_batch(gvl_lm* const* lms, int n_seq, int n_steps, long long* tokens_out, float* logits_out, long long eos_id,
                        long long pad_id, void* stream) {
    if (!lms || n_seq < 1 || n_seq > 4 || n_steps < 0) return GVL_ERR_ARG;
    gvl_lm* L0 = lms[0];
    if (!L0) return GVL_ERR_ARG;
    for (int b = 0; b < n_seq; ++b) {
        gvl_lm* m = lms[b];
        if (!m) return GVL_ERR_ARG;
        for (int c = 0; c < b; ++c)
            if (lms[c] == m) return GVL_ERR_ARG;                                   // one object = one sequence
        // same weights (the objects differ in KV cache, decode state, RoPE table and max_ctx only)
        if (m->w.n_layers != L0->w.n_layers || m->w.dim != L0->w.dim || m->w.vocab != L0->w.vocab || m->w.embed != L0->w.embed ||
            m->w.lm_head_w != L0->w.lm_head_w || m->layers[0].qkv_w != L0->layers[0].qkv_w) return GVL_ERR_ARG;
        if (m->host_ctx < 0 || m->host_ctx + n_steps > m->w.max_ctx) return GVL_ERR_STATE;
    }
    if (n_steps == 0) return GVL_OK;
    cudaStream_t caller = reinterpret_cast<cudaStream_t>(stream);
    cudaStream_t s = L0->cs;
    const gvl_lm_weights& w = L0->w;
    const int qkv_n = (w.heads + 2 * w.kv_heads) * w.head_dim;
    if (!L0->bx) {
        CK(dev_alloc(&L0->bx, (size_t)4 * w.dim));
        CK(dev_alloc(&L0->bqkv, (size_t)4 * qkv_n));
        CK(dev_alloc(&L0->bq, (size_t)4 * w.heads * w.head_dim));
        CK(dev_alloc(&L0->battn, (size_t)4 * w.heads * w.head_dim));
        CK(dev_alloc(&L0->bmid, (size_t)4 * w.ffn));
        CK(dev_alloc(&L0->blogits, (size_t)4 * w.vocab));
    }
    CU(cudaEventRecord(L0->ev_in, caller));
    CU(cudaStreamWaitEvent(s, L0->ev_in, 0));
    const bool want_logits = logits_out != nullptr;
    bool regraph = L0->bgraph == nullptr || L0->b_n != n_seq || L0->b_logits != want_logits || L0->b_eos != eos_id || L0->b_pad != pad_id;
    int zero = 0;
    for (int b = 0; b < n_seq; ++b) {
        gvl_lm* m = lms[b];
        if (L0->b_members[b] != m) regraph = true;
        if (want_logits && m->logit_steps < n_steps) {
            cudaFree(m->logit_buf);
            m->logit_buf = nullptr;
            CK(dev_alloc(&m->logit_buf, (size_t)n_steps * w.vocab));
            m->logit_steps = n_steps;
            if (m->graph) { cudaGraphExecDestroy(m->graph); m->graph = nullptr; }      // the single-sequence graph bound the old buffer
            regraph = true;
        }
        CU(cudaMemcpyAsync(&m->st->step, &zero, sizeof(int), cudaMemcpyHostToDevice, s));
    }
    if (regraph) {
        if (L0->bgraph) { cudaGraphExecDestroy(L0->bgraph); L0->bgraph = nullptr; }
        if (!(L0->b_warm_mask & (1u << n_seq))) {
            // one eager step outside capture (cudaFuncSetAttribute must not be issued while capturing); undone by restoring the states
            DecodeState saved[4];
            for (int b = 0; b < n_seq; ++b) CU(cudaMemcpyAsync(&saved[b], lms[b]->st, sizeof(DecodeState), cudaMemcpyDeviceToHost, s));
            CU(cudaStreamSynchronize(s));
            g_pdl = L0->use_pdl;
            int rc0 = enqueue_batch_step(lms, n_seq, -1, 0, false, s);
            g_pdl = false;
            if (rc0 != GVL_OK) return rc0;
            for (int b = 0; b < n_seq; ++b) CU(cudaMemcpyAsync(lms[b]->st, &saved[b], sizeof(DecodeState), cudaMemcpyHostToDevice, s));
            CU(cudaStreamSynchronize(s));
            L0->b_warm_mask |= 1u << n_seq;
        }
        cudaGraph_t g = nullptr;
        CU(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        const long long launches_before = g_launch_count;
        g_pdl = L0->use_pdl;
        int rc = enqueue_batch_step(lms, n_seq, eos_id, pad_id, want_logits, s);
        g_pdl = false;
        cudaError_t e = cudaStreamEndCapture(s, &g);
        g_launch_count = launches_before;
        if (rc != GVL_OK || e != cudaSuccess) { if (g) cudaGraphDestroy(g); return rc != GVL_OK ? rc : GVL_ERR_CUDA; }
        e = cudaGraphInstantiate(&L0->bgraph, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) return GVL_ERR_CUDA;
        L0->b_n = n_seq; L0->b_logits = want_logits; L0->b_eos = eos_id; L0->b_pad = pad_id;
        for (int b = 0; b < 4; ++b) L0->b_members[b] = b < n_seq ? lms[b] : nullptr;
    }
    const long long per_step = 2LL * n_seq + (long long)w.n_layers * (4 + 2) + 1 + n_seq;
    for (int i = 0; i < n_steps; ++i) {
        CU(cudaGraphLaunch(L0->bgraph, s));
        g_launch_count += per_step;
    }
    for (int b = 0; b < n_seq; ++b) {
        gvl_lm* m = lms[b];
        if (tokens_out) CU(cudaMemcpyAsync(tokens_out + (size_t)b * n_steps, m->tok_buf, (size_t)n_steps * sizeof(long long), cudaMemcpyDeviceToDevice, s));
        if (want_logits) CU(cudaMemcpyAsync(logits_out + (size_t)b * n_steps * w.vocab, m->logit_buf, (size_t)n_steps * w.vocab * sizeof(float),
                                            cudaMemcpyDeviceToDevice, s));
        m->host_ctx += n_steps;
    }
    CU(cudaEventRecord(L0->ev_out, s));
    CU(cudaStreamWaitEvent(caller, L0->ev_out, 0));
    return GVL_OK;
}

int gvl_lm_mega_trace(gvl_lm* lm, long long* host_out, int max_ctas, int* n_ctas, int* stride) {
    if (!lm || !host_out || !lm->trace) return GVL_ERR_STATE;
    const int n = num_sms() < max_ctas ? num_sms() : max_ctas;
    if (cudaDeviceSynchronize() != cudaSuccess) return GVL_ERR_CUDA;
    CU(cudaMemcpy(host_out, lm->trace, (size_t)n * MEGA_TRACE_STRIDE * sizeof(long long), cudaMemcpyDeviceToHost));
    if (n_ctas) *n_ctas = n;
    if (stride) *stride = MEGA_TRACE_STRIDE;
    return GVL_OK;
}

int gvl_lm_attention_split(int ctx, int heads, int n_ctas, int* warps_per_head, int* tokens_per_warp, int* max_partials, int* capacity) {
    if (ctx < 1 || heads < 1 || heads > 64 || n_ctas < 1 || n_ctas * 8 < heads || !warps_per_head || !tokens_per_warp || !max_partials ||
        !capacity)
        return GVL_ERR_ARG;
    decode_mega_attention_split(ctx, heads, n_ctas, warps_per_head, tokens_per_warp, max_partials);
    *capacity = n_ctas / heads + 2;             // MegaPlan::att_maxp (decode_mega_finalize)
    return GVL_OK;
}

int gvl_lm_decode_kind(const gvl_lm* lm) { return (lm && lm->use_mega && lm->plan_dev) ? 1 : 0; }

const long long* gvl_lm_first_token(gvl_lm* lm) { return lm ? lm->first_tok : nullptr; }

int gvl_lm_set_next_token(gvl_lm* lm, const long long* token_dev, void* stream) {
    if (!lm || !token_dev) return GVL_ERR_ARG;
    cudaStream_t caller = reinterpret_cast<cudaStream_t>(stream);
    // decode steps run on the object's own stream and fence themselves against the caller's with events (gvl_lm_decode), so a
    // copy enqueued on the caller's stream is ordered after the previous step and before the next one
    CU(cudaMemcpyAsync(&lm->st->cur_token, token_dev, sizeof(long long), cudaMemcpyDeviceToDevice, caller));
    return GVL_OK;
}

int gvl_lm_set_graph(gvl_lm* lm, int on) {
    if (!lm) return GVL_ERR_ARG;
    lm->use_graph = on != 0;
    return GVL_OK;
}

}  // extern "C"
