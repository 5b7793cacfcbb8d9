// Stage entry points for the two vision streams: host-side orchestration of the kernels in this
// library, one call per stage, all work enqueued on the caller's stream (no host sync inside).
//   gvl_clip_encode : CLIPVisionTransformer.forward up to hidden_states[-2]
//                     (modeling_clip.py:830-872, 578-657; consumer llava_next_video.py:504-505)
//   gvl_iv2_encode  : PretrainInternVideo2.forward(x, None, False, x_vis_return_idx=-2, x_vis_only=True)
//                     (internvideo2.py:970-1040; consumer llava_next_video.py:532)
#include "gvl_internal.h"
#include "../../include/gvl.h"

using namespace gvl;

namespace {

struct Carver {
    uint8_t* p;
    size_t off = 0;
    explicit Carver(void* base) : p(reinterpret_cast<uint8_t*>(base)) {}
    template <typename T>
    T* take(size_t n) {
        off = (off + 255) & ~size_t(255);
        T* r = p ? reinterpret_cast<T*>(p + off) : nullptr;
        off += n * sizeof(T);
        return r;
    }
};

struct ClipBufs {
    __nv_bfloat16 *col, *patch, *h, *qkv, *attn, *mid;
    float* x0;
    size_t bytes;
};
ClipBufs carve_clip(const gvl_clip_weights* w, int n_img, void* base) {
    Carver c(base);
    ClipBufs b;
    const size_t T = (size_t)n_img * (w->n_patch + 1);
    b.col = c.take<__nv_bfloat16>((size_t)n_img * w->n_patch * w->kpad);
    b.patch = c.take<__nv_bfloat16>((size_t)n_img * w->n_patch * w->dim);
    b.x0 = c.take<float>(T * w->dim);
    b.h = c.take<__nv_bfloat16>(T * w->dim);
    b.qkv = c.take<__nv_bfloat16>(T * 3 * w->dim);
    b.attn = c.take<__nv_bfloat16>(T * w->dim);
    b.mid = c.take<__nv_bfloat16>(T * w->ffn);
    b.bytes = c.off + 256;
    return b;
}

struct Iv2Bufs {
    __nv_bfloat16 *col, *patch, *h, *qkv, *attn, *mid;
    size_t bytes;
};
Iv2Bufs carve_iv2(const gvl_iv2_weights* w, int n_seg, void* base) {
    Carver c(base);
    Iv2Bufs b;
    const size_t np = (size_t)w->frames * 256;
    const size_t T = (size_t)n_seg * (np + 1);
    b.col = c.take<__nv_bfloat16>((size_t)n_seg * np * w->kpad);
    b.patch = c.take<__nv_bfloat16>((size_t)n_seg * np * w->dim);
    b.h = c.take<__nv_bfloat16>(T * w->dim);
    b.qkv = c.take<__nv_bfloat16>(T * 3 * w->heads * w->head_dim_pad);
    b.attn = c.take<__nv_bfloat16>(T * w->dim);
    b.mid = c.take<__nv_bfloat16>(T * w->ffn);
    b.bytes = c.off + 256;
    return b;
}

#define CK(expr)                 \
    do {                           \
        int _rc = (expr);          \
        if (_rc != GVL_OK) return _rc; \
    } while (0)

}  // namespace

extern "C" {

size_t gvl_clip_workspace(const gvl_clip_weights* w, int n_img) {
    if (!w || n_img <= 0) return 0;
    return carve_clip(w, n_img, nullptr).bytes;
}

int gvl_clip_encode(const gvl_clip_weights* w, const float* pix, float* hs, int n_img, void* workspace,
                    size_t ws_bytes, void* stream) {
    if (!w || !pix || !hs || !workspace || n_img <= 0) return GVL_ERR_ARG;
    if (w->image % 14 != 0 || (w->image / 14) * (w->image / 14) != w->n_patch) return GVL_ERR_ARG;
    if (w->dim % w->heads != 0) return GVL_ERR_ARG;
    ClipBufs b = carve_clip(w, n_img, workspace);
    if (b.bytes > ws_bytes) return GVL_ERR_NOMEM;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const int D = w->dim, F = w->ffn, H = w->heads, hd = D / H;
    const int S = w->n_patch + 1;
    const int T = n_img * S;
    float* x = hs;  // the fp32 residual stream lives in the caller's output buffer

    // embeddings (modeling_clip.py:182-191): conv-as-GEMM, cls, +pos; then pre_layrnorm (:851)
    CK(im2col_patch14(pix, 1, b.col, n_img, 3, 1, w->image, w->kpad, s));
    CK(gemm_bf16(b.col, w->kpad, w->patch_w, w->kpad, b.patch, D, n_img * w->n_patch, D, w->kpad, nullptr, nullptr,
                 nullptr, 0, GVL_ACT_NONE, GVL_RES_NONE, 0, 0, s));
    CK(clip_assemble(b.patch, (const float*)w->cls, (const float*)w->pos, b.x0, n_img, w->n_patch, D, s));
    CK(layernorm_f32_to_f32(b.x0, (const float*)w->pre_ln_w, (const float*)w->pre_ln_b, x, T, D, 1e-5f, s));

    for (int l = 0; l < w->n_layers; ++l) {
        const gvl_clip_layer& L = w->layers[l];
        // CLIPEncoderLayer.forward (modeling_clip.py:355-393)
        CK(layernorm_f32_to_bf16(x, (const float*)L.ln1_w, (const float*)L.ln1_b, b.h, T, D, 1e-5f, s));
        CK(gemm_bf16(b.h, D, L.qkv_w, D, b.qkv, 3 * D, T, 3 * D, D, L.qkv_b, nullptr, nullptr, 0, GVL_ACT_NONE,
                     GVL_RES_NONE, 0, 0, s));
        AttnArgs a;
        a.q = b.qkv; a.k = b.qkv + D; a.v = b.qkv + 2 * D; a.o = b.attn;
        a.q_bs = a.k_bs = a.v_bs = (long long)S * 3 * D;
        a.q_ts = a.k_ts = a.v_ts = 3 * D;
        a.q_hs = a.k_hs = a.v_hs = hd;
        a.o_bs = (long long)S * D; a.o_ts = D; a.o_hs = hd;
        a.batch = n_img; a.heads = H; a.kv_heads = H; a.sq = S; a.skv = S; a.head_dim = hd;
        a.scale = 1.0f;  // q rows of qkv_w are pre-scaled by hd^-0.5 (exact: power of two)
        a.causal = 0; a.round_scores = 1;
        CK(attention_fwd(a, s));
        CK(gemm_bf16(b.attn, D, L.out_w, D, x, D, T, D, D, L.out_b, nullptr, x, D, GVL_ACT_NONE, GVL_RES_F32, 1, 0, s));
        CK(layernorm_f32_to_bf16(x, (const float*)L.ln2_w, (const float*)L.ln2_b, b.h, T, D, 1e-5f, s));
        CK(gemm_bf16(b.h, D, L.fc1_w, D, b.mid, F, T, F, D, L.fc1_b, nullptr, nullptr, 0, GVL_ACT_QUICK_GELU,
                     GVL_RES_NONE, 0, 0, s));
        CK(gemm_bf16(b.mid, F, L.fc2_w, F, x, D, T, D, F, L.fc2_b, nullptr, x, D, GVL_ACT_NONE, GVL_RES_F32, 1, 0, s));
    }
    return GVL_OK;
}

size_t gvl_iv2_workspace(const gvl_iv2_weights* w, int n_seg) {
    if (!w || n_seg <= 0) return 0;
    return carve_iv2(w, n_seg, nullptr).bytes;
}

int gvl_iv2_encode(const gvl_iv2_weights* w, const float* pix, void* x_out, int n_seg, void* workspace,
                   size_t ws_bytes, void* stream) {
    if (!w || !pix || !x_out || !workspace || n_seg <= 0) return GVL_ERR_ARG;
    if (w->dim % w->heads != 0 || (w->dim / w->heads) % 8 != 0) return GVL_ERR_ARG;
    Iv2Bufs b = carve_iv2(w, n_seg, workspace);
    if (b.bytes > ws_bytes) return GVL_ERR_NOMEM;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const int D = w->dim, F = w->ffn, H = w->heads, hd = D / H;
    const int hdp = w->head_dim_pad, Dp = H * hdp;
    if (hdp < hd || hdp % 8 != 0 || hdp > 128) return GVL_ERR_ARG;
    const int np = w->frames * 256;
    const int S = np + 1;
    const int T = n_seg * S;
    __nv_bfloat16* x = reinterpret_cast<__nv_bfloat16*>(x_out);

    // PatchEmbed (internvideo2.py:721-725) + cls + pos_embed (:975-1005)
    CK(im2col_patch14(pix, 1, b.col, n_seg, 3, w->frames, 224, w->kpad, s));
    CK(gemm_bf16(b.col, w->kpad, w->patch_w, w->kpad, b.patch, D, n_seg * np, D, w->kpad, w->patch_b, nullptr, nullptr,
                 0, GVL_ACT_NONE, GVL_RES_NONE, 0, 0, s));
    CK(iv2_assemble(b.patch, (const __nv_bfloat16*)w->cls, (const __nv_bfloat16*)w->pos, x, n_seg, np, D, s));

    const float scale = 1.0f / sqrtf((float)hd);
    for (int l = 0; l < w->n_blocks; ++l) {
        const gvl_iv2_block& B = w->blocks[l];
        // Block._inner_forward (internvideo2.py:680-684): x += ls1(attn(norm1(x))); x += ls2(mlp(norm2(x)))
        CK(rmsnorm_bf16(x, D, (const __nv_bfloat16*)B.norm1_w, b.h, D, T, D, 1e-6f, s));
        // q/k/v leave the GEMM with every head zero-padded hd -> hdp (zero weight rows), see gvl.h
        CK(gemm_bf16(b.h, D, B.qkv_w, D, b.qkv, 3 * Dp, T, 3 * Dp, D, nullptr, nullptr, nullptr, 0, GVL_ACT_NONE,
                     GVL_RES_NONE, 0, 0, s));
        CK(iv2_qk_rmsnorm(b.qkv, (const __nv_bfloat16*)B.q_norm_w, (const __nv_bfloat16*)B.k_norm_w, T, Dp, 1e-6f, s, D));
        AttnArgs a;
        a.q = b.qkv; a.k = b.qkv + Dp; a.v = b.qkv + 2 * Dp; a.o = b.attn;
        a.q_bs = a.k_bs = a.v_bs = (long long)S * 3 * Dp;
        a.q_ts = a.k_ts = a.v_ts = 3 * Dp;
        a.q_hs = a.k_hs = a.v_hs = hdp;
        a.o_bs = (long long)S * D; a.o_ts = D; a.o_hs = hd;
        a.batch = n_seg; a.heads = H; a.kv_heads = H; a.sq = S; a.skv = S; a.head_dim = hdp; a.o_dim = hd;
        a.scale = scale; a.causal = 0; a.round_scores = 0;
        CK(attention_fwd(a, s));
        CK(gemm_bf16(b.attn, D, B.proj_w, D, x, D, T, D, D, B.proj_b, (const float*)B.ls1, x, D, GVL_ACT_NONE,
                     GVL_RES_BF16, 0, 0, s));
        CK(rmsnorm_bf16(x, D, (const __nv_bfloat16*)B.norm2_w, b.h, D, T, D, 1e-6f, s));
        CK(gemm_bf16(b.h, D, B.fc1_w, D, b.mid, F, T, F, D, B.fc1_b, nullptr, nullptr, 0, GVL_ACT_GELU_ERF, GVL_RES_NONE,
                     0, 0, s));
        CK(gemm_bf16(b.mid, F, B.fc2_w, F, x, D, T, D, F, B.fc2_b, (const float*)B.ls2, x, D, GVL_ACT_NONE, GVL_RES_BF16,
                     0, 0, s));
    }
    return GVL_OK;
}

}  // extern "C"
