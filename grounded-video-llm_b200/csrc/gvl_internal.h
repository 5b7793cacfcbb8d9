// Internal C++ declarations shared by the .cu translation units of libgvl.so.
// The public C ABI is include/gvl.h; everything here is namespace gvl.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stddef.h>

// status codes (mirrored in include/gvl.h)
#define GVL_OK 0
#define GVL_ERR_ARG (-1)
#define GVL_ERR_ALIGN (-2)
#define GVL_ERR_CUDA (-3)
#define GVL_ERR_DRIVER (-4)
#define GVL_ERR_NOMEM (-5)
#define GVL_ERR_STATE (-6)

namespace gvl {

extern long long g_launch_count;  // kernels launched by this library (bench.py: gpu_launches)

int num_sms();

// Per-kernel "function attributes already set" state is PER DEVICE (cudaFuncSetAttribute applies to the current device's copy of
// the function): returns true the first time it is called for `mask` on the current device. One 64-bit mask per call site.
inline bool first_use_on_device(unsigned long long& mask) {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (mask & bit) return false;
    mask |= bit;
    return true;
}

// Programmatic dependent launch (PDL) for the decode chain: when g_pdl is set, kernels are launched with
// programmaticStreamSerialization so that kernel N+1 is scheduled while kernel N drains; every such kernel calls
// pdl_wait() (griddepcontrol.wait) before it touches anything its predecessor wrote. Weight prefetches are issued
// before the wait, which hides launch latency and the first DRAM round trip of the weight stream.
extern bool g_pdl;
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    if (g_pdl) {
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// profile.cu -- optional CUDA-event timing per kernel family (off by default)
#define GVL_PROF_GEMM 0
#define GVL_PROF_ATTN 1
#define GVL_PROF_GEMV 2
#define GVL_PROF_DECODE_ATTN 3
#define GVL_PROF_KINDS 4
bool prof_enabled();
void prof_begin(int kind, double work, cudaStream_t s);
void prof_end(int kind, cudaStream_t s);

// gemm_tcgen05.cu
int make_tmap_2d_bf16(CUtensorMap* tm, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols);
int gemm_bf16(const void* A, int lda, const void* W, int ldw, void* out, int ldo, int M, int N, int K,
              const void* bias, const float* gamma, const void* residual, int ldr, int act, int res,
              int out_f32, int bn_hint, cudaStream_t stream);

// gemm_tcgen05_2cta.cu (cta_group::2, 256x256 tiles per CTA pair; same contract, N tile fixed at 256)
int gemm_bf16_2cta(const void* A, int lda, const void* W, int ldw, void* out, int ldo, int M, int N, int K,
                   const void* bias, const float* gamma, const void* residual, int ldr, int act, int res, int out_f32,
                   cudaStream_t stream);

// attention.cu  (q/k/v: bf16, element strides given per token and per head)
struct AttnArgs {
    const __nv_bfloat16* q;
    const __nv_bfloat16* k;
    const __nv_bfloat16* v;
    __nv_bfloat16* o;
    long long q_bs, q_ts, q_hs;  // batch / token / head strides (elements)
    long long k_bs, k_ts, k_hs;
    long long v_bs, v_ts, v_hs;
    long long o_bs, o_ts, o_hs;
    int batch, heads, kv_heads, sq, skv, head_dim;
    float scale;
    int causal;        // bottom-right aligned causal mask (flash-attn convention)
    int round_scores;  // 1: round q*k^T to bf16 before softmax (eager bmm path, CLIP)
    int o_dim = 0;     // columns of each head actually written to o (0 = head_dim); IV2 stores 88 of a padded 96
};
int attention_fwd(const AttnArgs& a, cudaStream_t stream);        // dispatch: tcgen05 kernel when the shape allows
int attention_mma_fwd(const AttnArgs& a, cudaStream_t stream);    // attention.cu  (mma.sync, any head_dim % 8 == 0)
bool attention_tc_supported(const AttnArgs& a);                   // attention_tc.cu (tcgen05/TMEM/TMA, d in 64/96/128)
int attention_tc_fwd(const AttnArgs& a, cudaStream_t stream);
int attention_tc2_fwd(const AttnArgs& a, cudaStream_t stream);    // attention_tc2.cu (two query tiles per CTA)

// rowops.cu
int layernorm_f32_to_bf16(const float* x, const float* w, const float* b, __nv_bfloat16* y, int rows,
                          int cols, float eps, cudaStream_t s);
int layernorm_f32_to_f32(const float* x, const float* w, const float* b, float* y, int rows, int cols, float eps,
                         cudaStream_t s);
int rmsnorm_bf16(const __nv_bfloat16* x, long long ldx, const __nv_bfloat16* w, __nv_bfloat16* y,
                 long long ldy, int rows, int cols, float eps, cudaStream_t s);
int iv2_qk_rmsnorm(__nv_bfloat16* qkv, const __nv_bfloat16* wq, const __nv_bfloat16* wk, int rows,
                   int dim, float eps, cudaStream_t s, int n_real = 0);

// gather.cu
int im2col_patch14(const void* pix, int pix_is_f32, __nv_bfloat16* out, int n_img, int chans, int frames,
                   int hw, int kpad, cudaStream_t s);
int clip_assemble(const __nv_bfloat16* patch, const float* cls, const float* pos, float* x, int n_img,
                  int n_patch, int dim, cudaStream_t s);
int iv2_assemble(const __nv_bfloat16* patch, const __nv_bfloat16* cls, const __nv_bfloat16* pos,
                 __nv_bfloat16* x, int n_seg, int n_patch, int dim, cudaStream_t s);
int hd_merge_newline(const float* hs, const float* sub_gn, __nv_bfloat16* out, int n_img, cudaStream_t s);
int iv2_pool(const __nv_bfloat16* x, __nv_bfloat16* out, int n_seg, int frames, int dim, cudaStream_t s);
int clip_pool3(const float* hs, __nv_bfloat16* out, int n_img, cudaStream_t s);
int visual_concat(const __nv_bfloat16* a, int a_rows, const __nv_bfloat16* b, int b_rows,
                  const __nv_bfloat16* newline, __nv_bfloat16* out, int n_seg, int dim, cudaStream_t s);
int embed_splice(const long long* ids, int t_text, int img_pos, const __nv_bfloat16* table,
                 const __nv_bfloat16* visual, int n_vis, __nv_bfloat16* out, int dim, int vis_last,
                 cudaStream_t s);
int rope_qkv_cache(const __nv_bfloat16* qkv, __nv_bfloat16* q_out, __nv_bfloat16* k_cache,
                   __nv_bfloat16* v_cache, const __nv_bfloat16* cos, const __nv_bfloat16* sin,
                   const int* positions, int tokens, int heads, int kv_heads, int head_dim, int pos0,
                   int max_ctx, cudaStream_t s);

}  // namespace gvl
