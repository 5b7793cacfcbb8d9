// Decode-side weight-streaming GEMV: y[M,N] = x[M,K] W[N,K]^T for M <= 4.
// HBM-bound: what matters is that every SM keeps >= ~64 KB of weight loads in flight ALL the time.
//   * one work unit = two weight rows (SwiGLU: the gate row and the up row of one output column); a warp walks its
//     units as a flat sequence of (unit, batch) items, a batch being U x 2 independent 128-bit streaming loads
//     (ld.global.nc.L1::no_allocate) per lane;
//   * register double buffering: the loads of item i+1 are issued BEFORE the FMAs of item i, so the memory system
//     never drains while a warp computes or reduces (the first-generation kernel alternated load / compute phases
//     in lock-step across all warps and reached only ~45 % of HBM bandwidth, profiles/r1_decode.md);
//   * the first batch is issued before the x staging / RMSNorm prologue (and before griddepcontrol.wait under PDL),
//     so the weight stream's first DRAM round trip overlaps the prologue and the tail of the previous kernel;
//   * x (optionally RMS-normalised with exactly the rounding of rmsnorm_bf16) lives in shared memory as bf16.
// Reference call sites: the q_len=1 passes of Phi3DecoderLayer / LlamaDecoderLayer (modeling_phi3.py:1034-1095,
// modeling_llama.py:699-760) and lm_head + .float() (modeling_phi3.py:1525-1526).
#include "gvl_internal.h"
#include "ptx.cuh"
#include "decode.h"

namespace gvl {

namespace {

constexpr int GV_THREADS = 256;
constexpr int GV_WARPS = GV_THREADS / 32;
constexpr int GV_MAXM = 4;
constexpr int GV_U = 4;          // 128-bit loads per lane per row per batch
constexpr int GV_CTAS_PER_SM = 2;

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ float dot8(uint4 w, uint4 x) {
    float2 a, b;
    float s;
    a = unpack_bf16(w.x); b = unpack_bf16(x.x); s = a.x * b.x + a.y * b.y;
    a = unpack_bf16(w.y); b = unpack_bf16(x.y); s += a.x * b.x + a.y * b.y;
    a = unpack_bf16(w.z); b = unpack_bf16(x.z); s += a.x * b.x + a.y * b.y;
    a = unpack_bf16(w.w); b = unpack_bf16(x.w); s += a.x * b.x + a.y * b.y;
    return s;
}

struct Batch {
    uint4 v0[GV_U], v1[GV_U];
};

template <int MT, bool SWIGLU>
__global__ void __launch_bounds__(GV_THREADS, GV_CTAS_PER_SM)
gemv3_kernel(const __nv_bfloat16* __restrict__ x, int ldx, const __nv_bfloat16* __restrict__ W, int ldw, void* out,
             int ldo, int N, int K, const __nv_bfloat16* __restrict__ norm_w, float eps,
             const __nv_bfloat16* __restrict__ bias, const __nv_bfloat16* residual /* may alias out */, int ldr,
             int out_f32) {
    extern __shared__ __align__(16) uint8_t smem[];
    __nv_bfloat16* sx = reinterpret_cast<__nv_bfloat16*>(smem);  // [MT][K]
    __shared__ float s_red[MT][GV_WARPS];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int kv = K / 8;                                  // 128-bit chunks per row
    const int n_units = SWIGLU ? N / 2 : (N + 1) / 2;
    const int gw = blockIdx.x * GV_WARPS + warp;
    const int nw = gridDim.x * GV_WARPS;
    const int nb = (kv + 32 * GV_U - 1) / (32 * GV_U);     // batches per unit
    const int my_units = gw < n_units ? (n_units - gw + nw - 1) / nw : 0;
    const int n_items = my_units * nb;

    auto rows_of = [&](int unit, int& r0, int& r1) {
        if (SWIGLU) {
            r0 = (unit / 128) * 256 + (unit % 128);        // gate row (interleaved per 256-row block, gvl/weights.py)
            r1 = r0 + 128;                                 // up row
        } else {
            r0 = unit * 2;
            r1 = (r0 + 1 < N) ? r0 + 1 : r0;
        }
    };
    auto load_item = [&](int item, Batch& bt) {
        const int unit = gw + (item / nb) * nw, bidx = item % nb;
        int r0, r1;
        rows_of(unit, r0, r1);
        const uint4* w0 = reinterpret_cast<const uint4*>(W + (size_t)r0 * ldw);
        const uint4* w1 = reinterpret_cast<const uint4*>(W + (size_t)r1 * ldw);
#pragma unroll
        for (int u = 0; u < GV_U; ++u) {
            const int i = (bidx * GV_U + u) * 32 + lane;
            if (i < kv) { bt.v0[u] = ldg_stream(w0 + i); bt.v1[u] = ldg_stream(w1 + i); }
        }
    };

    // ---- weights first: their DRAM latency overlaps the prologue (and, under PDL, the previous kernel's tail)
    Batch bufA, bufB;
    if (n_items > 0) load_item(0, bufA);
    pdl_launch_dependents();
    pdl_wait();

    // ---- stage x (with optional RMSNorm, rounding identical to rmsnorm_bf16)
    float ss[MT];
#pragma unroll
    for (int m = 0; m < MT; ++m) ss[m] = 0.f;
    for (int i = tid; i < kv; i += GV_THREADS) {
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            uint4 v = *(reinterpret_cast<const uint4*>(x + (size_t)m * ldx) + i);
            reinterpret_cast<uint4*>(sx + (size_t)m * K)[i] = v;
            if (norm_w != nullptr) {
                float2 f;
                f = unpack_bf16(v.x); ss[m] += f.x * f.x + f.y * f.y;
                f = unpack_bf16(v.y); ss[m] += f.x * f.x + f.y * f.y;
                f = unpack_bf16(v.z); ss[m] += f.x * f.x + f.y * f.y;
                f = unpack_bf16(v.w); ss[m] += f.x * f.x + f.y * f.y;
            }
        }
    }
    if (norm_w != nullptr) {
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            float v = wsum(ss[m]);
            if (lane == 0) s_red[m][warp] = v;
        }
        __syncthreads();
        float rstd[MT];
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < GV_WARPS; ++w) t += s_red[m][w];
            rstd[m] = rsqrtf(t / K + eps);
        }
        // each thread re-normalises exactly the chunks it staged itself: no barrier needed in between
        for (int i = tid; i < kv; i += GV_THREADS) {
            uint4 wv = __ldg(reinterpret_cast<const uint4*>(norm_w) + i);
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                uint4 v = reinterpret_cast<uint4*>(sx + (size_t)m * K)[i], o;
                float2 f, g;
                f = unpack_bf16(v.x); g = unpack_bf16(wv.x); o.x = pack_bf16(bf16r(f.x * rstd[m]) * g.x, bf16r(f.y * rstd[m]) * g.y);
                f = unpack_bf16(v.y); g = unpack_bf16(wv.y); o.y = pack_bf16(bf16r(f.x * rstd[m]) * g.x, bf16r(f.y * rstd[m]) * g.y);
                f = unpack_bf16(v.z); g = unpack_bf16(wv.z); o.z = pack_bf16(bf16r(f.x * rstd[m]) * g.x, bf16r(f.y * rstd[m]) * g.y);
                f = unpack_bf16(v.w); g = unpack_bf16(wv.w); o.w = pack_bf16(bf16r(f.x * rstd[m]) * g.x, bf16r(f.y * rstd[m]) * g.y);
                reinterpret_cast<uint4*>(sx + (size_t)m * K)[i] = o;
            }
        }
    }
    __syncthreads();

    float a0[MT], a1[MT];
#pragma unroll
    for (int m = 0; m < MT; ++m) { a0[m] = 0.f; a1[m] = 0.f; }

    auto consume = [&](int item, const Batch& bt) {
        const int bidx = item % nb;
#pragma unroll
        for (int u = 0; u < GV_U; ++u) {
            const int i = (bidx * GV_U + u) * 32 + lane;
            if (i < kv) {
#pragma unroll
                for (int m = 0; m < MT; ++m) {
                    const uint4 xv = reinterpret_cast<const uint4*>(sx + (size_t)m * K)[i];
                    a0[m] += dot8(bt.v0[u], xv);
                    a1[m] += dot8(bt.v1[u], xv);
                }
            }
        }
        if (bidx != nb - 1) return;
        // ---- unit finished: reduce across the warp, fused epilogue, store
        const int unit = gw + (item / nb) * nw;
#pragma unroll
        for (int m = 0; m < MT; ++m) { a0[m] = wsum(a0[m]); a1[m] = wsum(a1[m]); }
        if (lane == 0) {
            int r0, r1;
            rows_of(unit, r0, r1);
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                if (SWIGLU) {
                    // Phi3MLP: up * silu(gate) on bf16 tensors (modeling_phi3.py:461-462)
                    const float g = bf16r(a0[m]), u = bf16r(a1[m]);
                    reinterpret_cast<__nv_bfloat16*>(out)[(size_t)m * ldo + unit] = __float2bfloat16_rn(u * bf16r(silu_f(g)));
                } else {
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int oc = r0 + j;
                        if (oc >= N) continue;
                        float y = j == 0 ? a0[m] : a1[m];
                        if (bias) y += __bfloat162float(bias[oc]);
                        y = bf16r(y);
                        if (residual) y = bf16r(y + __bfloat162float(residual[(size_t)m * ldr + oc]));
                        if (out_f32) reinterpret_cast<float*>(out)[(size_t)m * ldo + oc] = y;
                        else reinterpret_cast<__nv_bfloat16*>(out)[(size_t)m * ldo + oc] = __float2bfloat16_rn(y);
                    }
                }
            }
        }
#pragma unroll
        for (int m = 0; m < MT; ++m) { a0[m] = 0.f; a1[m] = 0.f; }
    };

    // ---- flat (unit, batch) item stream, register double-buffered
    for (int item = 0; item < n_items; item += 2) {
        if (item + 1 < n_items) load_item(item + 1, bufB);
        consume(item, bufA);
        if (item + 1 >= n_items) break;
        if (item + 2 < n_items) load_item(item + 2, bufA);
        consume(item + 1, bufB);
    }
}

}  // namespace

int gemv_bf16(const __nv_bfloat16* x, int ldx, const __nv_bfloat16* W, int ldw, void* out, int ldo, int M, int N,
              int K, const __nv_bfloat16* norm_w, float eps, const __nv_bfloat16* bias,
              const __nv_bfloat16* residual, int ldr, int act, int out_f32, cudaStream_t s) {
    if (M < 1 || M > GV_MAXM || K % 256 != 0 || (act != 0 && act != 3)) return GVL_ERR_ARG;
    if (act == 3 && N % 256 != 0) return GVL_ERR_ARG;
    const size_t smem = (size_t)M * K * 2;
    prof_begin(GVL_PROF_GEMV, 2.0 * (double)N * K, s);  // algorithmic bytes: the weight matrix, read once
    const int units = act == 3 ? N / 2 : (N + 1) / 2;
    int grid = (units + GV_WARPS - 1) / GV_WARPS;
    const int cap = num_sms() * GV_CTAS_PER_SM;
    if (grid > cap) grid = cap;
#define GV_LAUNCH(MT, SW)                                                                                             \
    do {                                                                                                              \
        auto kern = gemv3_kernel<MT, SW>;                                                                             \
        static size_t max_set[64] = {};                        /* per device: cudaFuncSetAttribute is per device */          \
        int dev_ = 0;                                                                                                 \
        cudaGetDevice(&dev_);                                                                                         \
        if (smem + 1024 > 48 * 1024 && smem > max_set[dev_ & 63]) {   /* + the kernel's static shared memory (M = 3, K = 8192 is 48 KB + 96 B) */                                                          \
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)   \
                return GVL_ERR_CUDA;                                                                                  \
            max_set[dev_ & 63] = smem;                                                                                \
        }                                                                                                             \
        if (launch_k(kern, dim3(grid), dim3(GV_THREADS), smem, s, x, ldx, W, ldw, out, ldo, N, K, norm_w, eps, bias,    \
                     residual, ldr, out_f32) != cudaSuccess) return GVL_ERR_CUDA;                                       \
    } while (0)
    if (act == 3) {
        switch (M) { case 1: GV_LAUNCH(1, true); break; case 2: GV_LAUNCH(2, true); break;
                     case 3: GV_LAUNCH(3, true); break; default: GV_LAUNCH(4, true); break; }
    } else {
        switch (M) { case 1: GV_LAUNCH(1, false); break; case 2: GV_LAUNCH(2, false); break;
                     case 3: GV_LAUNCH(3, false); break; default: GV_LAUNCH(4, false); break; }
    }
#undef GV_LAUNCH
    prof_end(GVL_PROF_GEMV, s);
    g_launch_count++;
    return cudaGetLastError() == cudaSuccess ? GVL_OK : GVL_ERR_CUDA;
}

}  // namespace gvl
