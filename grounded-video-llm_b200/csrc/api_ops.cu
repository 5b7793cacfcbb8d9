// extern "C" operator-level entry points of libgvl.so (declared in include/gvl.h).
#include "gvl_internal.h"
#include "decode.h"
#include "../../include/gvl.h"

using namespace gvl;

static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" {

const char* gvl_version(void) { return "gvl-b200 0.1 (sm_100a)"; }
long long gvl_launch_count(void) { return g_launch_count; }

int gvl_gemm_bf16(const void* A, int lda, const void* W, int ldw, void* out, int ldo, int M, int N, int K,
                  const void* bias, const float* gamma, const void* residual, int ldr, int act, int res,
                  int out_f32, int bn_hint, void* stream) {
    if (!A || !W || !out) return GVL_ERR_ARG;
    if (res != GVL_RES_NONE && !residual) return GVL_ERR_ARG;
    return gemm_bf16(A, lda, W, ldw, out, ldo, M, N, K, bias, gamma, residual, ldr, act, res, out_f32, bn_hint,
                     S(stream));
}

int gvl_attention(const void* q, const void* k, const void* v, void* o, const long long* qs, const long long* ks,
                  const long long* vs, const long long* os, int batch, int heads, int kv_heads, int sq, int skv,
                  int head_dim, float scale, int causal, int round_scores, int o_dim, void* stream) {
    if (!q || !k || !v || !o || !qs || !ks || !vs || !os) return GVL_ERR_ARG;
    AttnArgs a;
    a.q = (const __nv_bfloat16*)q; a.k = (const __nv_bfloat16*)k; a.v = (const __nv_bfloat16*)v;
    a.o = (__nv_bfloat16*)o;
    a.q_bs = qs[0]; a.q_ts = qs[1]; a.q_hs = qs[2];
    a.k_bs = ks[0]; a.k_ts = ks[1]; a.k_hs = ks[2];
    a.v_bs = vs[0]; a.v_ts = vs[1]; a.v_hs = vs[2];
    a.o_bs = os[0]; a.o_ts = os[1]; a.o_hs = os[2];
    a.batch = batch; a.heads = heads; a.kv_heads = kv_heads; a.sq = sq; a.skv = skv; a.head_dim = head_dim;
    a.scale = scale; a.causal = causal; a.round_scores = round_scores; a.o_dim = o_dim;
    return attention_fwd(a, S(stream));
}

int gvl_layernorm_f32(const float* x, const float* w, const float* b, void* y, int rows, int cols, float eps,
                      void* stream) {
    return layernorm_f32_to_bf16(x, w, b, (__nv_bfloat16*)y, rows, cols, eps, S(stream));
}
int gvl_rmsnorm_bf16(const void* x, long long ldx, const void* w, void* y, long long ldy, int rows, int cols,
                     float eps, void* stream) {
    return rmsnorm_bf16((const __nv_bfloat16*)x, ldx, (const __nv_bfloat16*)w, (__nv_bfloat16*)y, ldy, rows, cols,
                        eps, S(stream));
}
int gvl_iv2_qk_rmsnorm(void* qkv, const void* wq, const void* wk, int rows, int dim, float eps, void* stream) {
    return iv2_qk_rmsnorm((__nv_bfloat16*)qkv, (const __nv_bfloat16*)wq, (const __nv_bfloat16*)wk, rows, dim, eps,
                          S(stream));
}
int gvl_im2col_patch14(const void* pix, int pix_is_f32, void* out, int n_img, int chans, int frames, int hw,
                       int kpad, void* stream) {
    return im2col_patch14(pix, pix_is_f32, (__nv_bfloat16*)out, n_img, chans, frames, hw, kpad, S(stream));
}
int gvl_clip_assemble(const void* patch, const float* cls, const float* pos, float* x, int n_img, int n_patch,
                      int dim, void* stream) {
    return clip_assemble((const __nv_bfloat16*)patch, cls, pos, x, n_img, n_patch, dim, S(stream));
}
int gvl_iv2_assemble(const void* patch, const void* cls, const void* pos, void* x, int n_seg, int n_patch, int dim,
                     void* stream) {
    return iv2_assemble((const __nv_bfloat16*)patch, (const __nv_bfloat16*)cls, (const __nv_bfloat16*)pos,
                        (__nv_bfloat16*)x, n_seg, n_patch, dim, S(stream));
}
int gvl_hd_merge_newline(const float* hs, const float* sub_gn, void* out, int n_img, void* stream) {
    return hd_merge_newline(hs, sub_gn, (__nv_bfloat16*)out, n_img, S(stream));
}
int gvl_iv2_pool(const void* x, void* out, int n_seg, int frames, int dim, void* stream) {
    return iv2_pool((const __nv_bfloat16*)x, (__nv_bfloat16*)out, n_seg, frames, dim, S(stream));
}
int gvl_clip_pool3(const float* hs, void* out, int n_img, void* stream) {
    return clip_pool3(hs, (__nv_bfloat16*)out, n_img, S(stream));
}
int gvl_embed_splice(const long long* ids, int t_text, int img_pos, const void* table, const void* visual,
                     int n_vis, void* out, int dim, int vis_last, void* stream) {
    return embed_splice(ids, t_text, img_pos, (const __nv_bfloat16*)table, (const __nv_bfloat16*)visual, n_vis,
                        (__nv_bfloat16*)out, dim, vis_last, S(stream));
}
int gvl_rope_qkv_cache(const void* qkv, void* q_out, void* k_cache, void* v_cache, const void* cosb,
                       const void* sinb, const int* positions, int tokens, int heads, int kv_heads, int head_dim,
                       int pos0, int max_ctx, void* stream) {
    return rope_qkv_cache((const __nv_bfloat16*)qkv, (__nv_bfloat16*)q_out, (__nv_bfloat16*)k_cache,
                          (__nv_bfloat16*)v_cache, (const __nv_bfloat16*)cosb, (const __nv_bfloat16*)sinb, positions,
                          tokens, heads, kv_heads, head_dim, pos0, max_ctx, S(stream));
}

int gvl_gemv_bf16(const void* x, int ldx, const void* W, int ldw, void* out, int ldo, int M, int N, int K,
                  const void* norm_w, float norm_eps, const void* bias, const void* residual, int ldr, int act,
                  int out_f32, void* stream) {
    if (!x || !W || !out) return GVL_ERR_ARG;
    return gemv_bf16((const __nv_bfloat16*)x, ldx, (const __nv_bfloat16*)W, ldw, out, ldo, M, N, K,
                     (const __nv_bfloat16*)norm_w, norm_eps, (const __nv_bfloat16*)bias,
                     (const __nv_bfloat16*)residual, ldr, act, out_f32, S(stream));
}
size_t gvl_decode_attention_workspace(int heads, int head_dim, int max_ctx) {
    return decode_attention_workspace(heads, head_dim, max_ctx);
}
int gvl_decode_attention(const void* q, const void* k_cache, const void* v_cache, void* o, float* workspace,
                         const int* ctx_len_dev, int heads, int kv_heads, int head_dim, int max_ctx, float scale,
                         void* stream) {
    if (!q || !k_cache || !v_cache || !o || !workspace || !ctx_len_dev) return GVL_ERR_ARG;
    return decode_attention((const __nv_bfloat16*)q, (const __nv_bfloat16*)k_cache, (const __nv_bfloat16*)v_cache,
                            (__nv_bfloat16*)o, workspace, ctx_len_dev, heads, kv_heads, head_dim, max_ctx, scale,
                            S(stream));
}
int gvl_argmax_f32(const float* logits, int n, long long* out_token, void* stream) {
    if (!logits || !out_token || n <= 0) return GVL_ERR_ARG;
    return argmax_f32(logits, n, out_token, S(stream));
}
int gvl_visual_concat(const void* a, int a_rows, const void* b, int b_rows, const void* newline, void* out, int n_seg,
                      int dim, void* stream) {
    return visual_concat((const __nv_bfloat16*)a, a_rows, (const __nv_bfloat16*)b, b_rows,
                         (const __nv_bfloat16*)newline, (__nv_bfloat16*)out, n_seg, dim, S(stream));
}
int gvl_layernorm_f32_out_f32(const float* x, const float* w, const float* b, float* y, int rows, int cols, float eps,
                              void* stream) {
    return layernorm_f32_to_f32(x, w, b, y, rows, cols, eps, S(stream));
}

}  // extern "C"
