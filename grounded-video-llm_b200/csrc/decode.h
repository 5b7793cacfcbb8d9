// Device-resident decode bookkeeping shared by decode.cu and model_lm.cu.
#pragma once
#include "gvl_internal.h"

namespace gvl {

struct DecodeState {
    int ctx_len;          // tokens in the KV cache == position id of the token processed next
    int attn_len;         // ctx_len + 1 while a step is in flight (what decode attention must cover)
    int step;             // index into tokens_out
    int finished;         // EOS seen
    long long cur_token;  // token to embed at the next step
};

int gemv_bf16(const __nv_bfloat16* x, int ldx, const __nv_bfloat16* W, int ldw, void* out, int ldo, int M, int N,
              int K, const __nv_bfloat16* norm_w, float eps, const __nv_bfloat16* bias,
              const __nv_bfloat16* residual, int ldr, int act, int out_f32, cudaStream_t s);
size_t decode_attention_workspace(int heads, int head_dim, int max_ctx);
int decode_attention(const __nv_bfloat16* q, const __nv_bfloat16* kc, const __nv_bfloat16* vc, __nv_bfloat16* o,
                     float* ws, const int* ctx_len_dev, int heads, int kv_heads, int head_dim, int max_ctx,
                     float scale, cudaStream_t s);
// batched forms for gvl_lm_decode_batch: one launch per layer for all sequences (each with its own cache / state / workspace)
struct DecodeAttnSeq {
    const __nv_bfloat16 *q, *kc, *vc;
    __nv_bfloat16* o;
    float* ws;              // decode_attention_workspace(heads, head_dim, max_ctx) of THIS sequence
    const int* ctx_len;
    int max_ctx;
};
struct DecodeAttnBatch { DecodeAttnSeq s[4]; };
struct DecodeRopeSeq {
    const __nv_bfloat16* qkv;
    __nv_bfloat16 *q_out, *k_cache, *v_cache;
    const __nv_bfloat16 *cosb, *sinb;
    const DecodeState* st;
    int max_ctx;
};
struct DecodeRopeBatch { DecodeRopeSeq s[4]; };
int decode_attention_batch(const DecodeAttnBatch& b, int n_seq, int heads, int kv_heads, int head_dim, float scale, cudaStream_t s);
int rope_decode_batch(const DecodeRopeBatch& b, int n_seq, int heads, int kv_heads, int D, cudaStream_t s);
int argmax_f32(const float* logits, int n, long long* out, cudaStream_t s);
int embed_token(const __nv_bfloat16* table, const DecodeState* st, __nv_bfloat16* x, int dim, int vocab, cudaStream_t s);
int rope_decode(const __nv_bfloat16* qkv, __nv_bfloat16* q_out, __nv_bfloat16* k_cache, __nv_bfloat16* v_cache,
                const __nv_bfloat16* cosb, const __nv_bfloat16* sinb, const DecodeState* st, int heads, int kv_heads,
                int D, int max_ctx, cudaStream_t s);
int step_begin(DecodeState* st, cudaStream_t s);
int step_end(const float* logits, int n, DecodeState* st, long long* tokens_out, float* logits_out, long long eos_id,
             long long pad_id, cudaStream_t s);

}  // namespace gvl
