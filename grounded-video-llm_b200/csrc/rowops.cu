// Row-wise normalisation kernels (HBM-bound; one warp per row, 128-bit accesses, shuffle reductions).
//   layernorm_f32_to_bf16 : CLIP LayerNorm on the fp32 residual stream (modeling_clip.py:351-353,
//                           824-826, eps 1e-5), output rounded to bf16 = the autocast cast at the
//                           following nn.Linear.
//   rmsnorm_bf16          : InternVideo2 RMSNorm (internvideo2.py:437-448), Phi3RMSNorm
//                           (modeling_phi3.py:310-324), LlamaRMSNorm (modeling_llama.py:74-88):
//                           fp32 mean-square, x*rsqrt rounded to bf16 BEFORE the bf16 weight multiply.
//   iv2_qk_rmsnorm        : the q_norm / k_norm over the flattened 1408-wide q and k rows
//                           (internvideo2.py:590-598), in place inside the packed qkv buffer.
#include "gvl_internal.h"
#include "ptx.cuh"

namespace gvl {

long long g_launch_count = 0;
bool g_pdl = false;

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

constexpr int ROW_WARPS = 4;

template <bool OUT_F32>
__global__ void __launch_bounds__(ROW_WARPS * 32)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                 void* __restrict__ yv, int rows, int cols, float eps) {
    const int row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float4* xr = reinterpret_cast<const float4*>(x + size_t(row) * cols);
    const int nv = cols / 4;
    float s = 0.f;
    for (int i = lane; i < nv; i += 32) {
        float4 v = xr[i];
        s += (v.x + v.y) + (v.z + v.w);
    }
    const float mean = warp_sum(s) / cols;
    float q = 0.f;
    for (int i = lane; i < nv; i += 32) {
        float4 v = xr[i];
        float a0 = v.x - mean, a1 = v.y - mean, a2 = v.z - mean, a3 = v.w - mean;
        q += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
    }
    const float rstd = rsqrtf(warp_sum(q) / cols + eps);
    const float4* w4 = reinterpret_cast<const float4*>(w);
    const float4* b4 = reinterpret_cast<const float4*>(b);
    for (int i = lane; i < nv; i += 32) {
        float4 v = xr[i], ww = __ldg(w4 + i), bb = __ldg(b4 + i);
        float4 r = make_float4((v.x - mean) * rstd * ww.x + bb.x, (v.y - mean) * rstd * ww.y + bb.y,
                               (v.z - mean) * rstd * ww.z + bb.z, (v.w - mean) * rstd * ww.w + bb.w);
        if (OUT_F32) {
            reinterpret_cast<float4*>(reinterpret_cast<float*>(yv) + size_t(row) * cols)[i] = r;
        } else {
            uint2 o;
            o.x = pack_bf16(r.x, r.y);
            o.y = pack_bf16(r.z, r.w);
            reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(yv) + size_t(row) * cols)[i] = o;
        }
    }
}

__device__ __forceinline__ void rms_row(const __nv_bfloat16* xr, const __nv_bfloat16* __restrict__ w,
                                        __nv_bfloat16* yr, int cols, float eps, int lane, int n_mean = 0) {
    const uint4* x4 = reinterpret_cast<const uint4*>(xr);
    const int nv = cols / 8;
    float ss = 0.f;
    for (int i = lane; i < nv; i += 32) {
        uint4 v = x4[i];
        float2 f;
        f = unpack_bf16(v.x); ss += f.x * f.x + f.y * f.y;
        f = unpack_bf16(v.y); ss += f.x * f.x + f.y * f.y;
        f = unpack_bf16(v.z); ss += f.x * f.x + f.y * f.y;
        f = unpack_bf16(v.w); ss += f.x * f.x + f.y * f.y;
    }
    // n_mean > 0: the row is zero-padded (IV2 heads stored 88 -> 96) and the mean runs over the real elements
    const float rstd = rsqrtf(warp_sum(ss) / (n_mean > 0 ? n_mean : cols) + eps);
    const uint4* w4 = reinterpret_cast<const uint4*>(w);
    uint4* y4 = reinterpret_cast<uint4*>(yr);
    for (int i = lane; i < nv; i += 32) {
        uint4 v = x4[i], ww = __ldg(w4 + i), o;
        float2 f, g;
        f = unpack_bf16(v.x); g = unpack_bf16(ww.x); o.x = pack_bf16(bf16r(f.x * rstd) * g.x, bf16r(f.y * rstd) * g.y);
        f = unpack_bf16(v.y); g = unpack_bf16(ww.y); o.y = pack_bf16(bf16r(f.x * rstd) * g.x, bf16r(f.y * rstd) * g.y);
        f = unpack_bf16(v.z); g = unpack_bf16(ww.z); o.z = pack_bf16(bf16r(f.x * rstd) * g.x, bf16r(f.y * rstd) * g.y);
        f = unpack_bf16(v.w); g = unpack_bf16(ww.w); o.w = pack_bf16(bf16r(f.x * rstd) * g.x, bf16r(f.y * rstd) * g.y);
        y4[i] = o;
    }
}

__global__ void __launch_bounds__(ROW_WARPS * 32)
rmsnorm_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, const __nv_bfloat16* __restrict__ w,
               __nv_bfloat16* __restrict__ y, long long ldy, int rows, int cols, float eps) {
    const int row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
    if (row >= rows) return;
    rms_row(x + row * ldx, w, y + row * ldy, cols, eps, threadIdx.x & 31);
}

__global__ void __launch_bounds__(ROW_WARPS * 32)
qk_rmsnorm_kernel(__nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ wq,
                  const __nv_bfloat16* __restrict__ wk, int rows, int dim, float eps, int n_real) {
    const int row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int part = blockIdx.y;  // 0 = q, 1 = k
    __nv_bfloat16* p = qkv + size_t(row) * 3 * dim + size_t(part) * dim;
    rms_row(p, part == 0 ? wq : wk, p, dim, eps, threadIdx.x & 31, n_real);
}

}  // namespace

int layernorm_f32_to_bf16(const float* x, const float* w, const float* b, __nv_bfloat16* y, int rows,
                          int cols, float eps, cudaStream_t s) {
    if (cols % 4 != 0 || rows <= 0) return GVL_ERR_ARG;
    layernorm_kernel<false><<<(rows + ROW_WARPS - 1) / ROW_WARPS, ROW_WARPS * 32, 0, s>>>(x, w, b, y, rows, cols, eps);
    g_launch_count++;
    return cudaGetLastError() == cudaSuccess ? GVL_OK : GVL_ERR_CUDA;
}

// in-place safe (each warp reads its whole row before the write pass touches it... it re-reads, so
// in == out is NOT allowed); used for CLIP pre_layrnorm (modeling_clip.py:851) whose output stays fp32.
int layernorm_f32_to_f32(const float* x, const float* w, const float* b, float* y, int rows, int cols, float eps,
                         cudaStream_t s) {
    if (cols % 4 != 0 || rows <= 0 || x == y) return GVL_ERR_ARG;
    layernorm_kernel<true><<<(rows + ROW_WARPS - 1) / ROW_WARPS, ROW_WARPS * 32, 0, s>>>(x, w, b, y, rows, cols, eps);
    g_launch_count++;
    return cudaGetLastError() == cudaSuccess ? GVL_OK : GVL_ERR_CUDA;
}

int rmsnorm_bf16(const __nv_bfloat16* x, long long ldx, const __nv_bfloat16* w, __nv_bfloat16* y,
                 long long ldy, int rows, int cols, float eps, cudaStream_t s) {
    if (cols % 8 != 0 || ldx % 8 != 0 || ldy % 8 != 0 || rows <= 0) return GVL_ERR_ARG;
    rmsnorm_kernel<<<(rows + ROW_WARPS - 1) / ROW_WARPS, ROW_WARPS * 32, 0, s>>>(x, ldx, w, y, ldy, rows, cols, eps);
    g_launch_count++;
    return cudaGetLastError() == cudaSuccess ? GVL_OK : GVL_ERR_CUDA;
}

int iv2_qk_rmsnorm(__nv_bfloat16* qkv, const __nv_bfloat16* wq, const __nv_bfloat16* wk, int rows, int dim,
                   float eps, cudaStream_t s, int n_real) {
    if (dim % 8 != 0 || rows <= 0) return GVL_ERR_ARG;
    dim3 grid((rows + ROW_WARPS - 1) / ROW_WARPS, 2);
    qk_rmsnorm_kernel<<<grid, ROW_WARPS * 32, 0, s>>>(qkv, wq, wk, rows, dim, eps, n_real);
    g_launch_count++;
    return cudaGetLastError() == cudaSuccess ? GVL_OK : GVL_ERR_CUDA;
}

}  // namespace gvl
