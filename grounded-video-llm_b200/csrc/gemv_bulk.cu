// Decode-side weight-streaming GEMV, fourth generation: the weight stream is staged through SHARED MEMORY with
// cp.async.bulk (the 1-D TMA path, SASS UBLKCP) instead of through registers.
//
// Why: a register-staged GEMV can keep at most (threads x free registers) bytes in flight per SM (~128 KB at 512
// threads), and only while every warp has work; probe numbers were 35-68 % of the HBM peak for the four per-layer
// GEMVs (profiles/r1_decode.md). Here one producer warp keeps a ring of weight-row segments (up to ~190 KB per SM)
// in flight regardless of what the consumer warps are doing.
//
//   grid = #SMs, 288 threads: warp 8 = producer (one elected lane issues cp.async.bulk, mbarrier complete_tx),
//   warps 0..7 = consumers. Stage s of the ring holds 8 segments, one per consumer warp; a segment is a contiguous
//   <= 4096-element piece of one weight row. Consumer warp w of CTA c owns the global "lane" c*8+w and walks its
//   units (1 row, or gate+up row pair for SwiGLU) segment by segment; all 8 lanes of a CTA advance one item per stage.
//   x (optionally RMS-normalised exactly like rmsnorm_bf16) is staged once per CTA in shared memory.
#include "gvl_internal.h"
#include "ptx.cuh"
#include "decode.h"

namespace gvl {

namespace {

constexpr int GB_CONSUMERS = 8;
constexpr int GB_THREADS = 32 * (GB_CONSUMERS + 1);
constexpr int GB_MAXM = 4;
constexpr int GB_SEG_MAX = 4096;      // elements per segment
constexpr int GB_SMEM_BUDGET = 200 * 1024;

__device__ __forceinline__ float wsum_b(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float dot8b(uint4 w, uint4 x) {
    float2 a, b;
    float s;
    a = unpack_bf16(w.x); b = unpack_bf16(x.x); s = a.x * b.x + a.y * b.y;
    a = unpack_bf16(w.y); b = unpack_bf16(x.y); s += a.x * b.x + a.y * b.y;
    a = unpack_bf16(w.z); b = unpack_bf16(x.z); s += a.x * b.x + a.y * b.y;
    a = unpack_bf16(w.w); b = unpack_bf16(x.w); s += a.x * b.x + a.y * b.y;
    return s;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
                 : "memory");
}

struct GbParams {
    const __nv_bfloat16* x; int ldx;
    const __nv_bfloat16* W; int ldw;
    void* out; int ldo;
    int N, K;
    const __nv_bfloat16* norm_w; float eps;
    const __nv_bfloat16* bias;
    const __nv_bfloat16* residual; int ldr;   // may alias out
    int out_f32;
    int nseg, seg_len, stages;                 // K = nseg * seg_len
};

template <int MT, bool SWIGLU>
__global__ void __launch_bounds__(GB_THREADS, 1)
gemv_bulk_kernel(const GbParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int K = p.K;
    const uint32_t x_bytes = (uint32_t)MT * K * 2;
    const uint32_t seg_bytes = (uint32_t)p.seg_len * 2;
    const uint32_t stage_bytes = GB_CONSUMERS * seg_bytes;
    __nv_bfloat16* sx = reinterpret_cast<__nv_bfloat16*>(smem);
    uint8_t* ring = smem + ((x_bytes + 127u) & ~127u);
    const uint32_t ring_u32 = ptx::smem_u32(ring);
    __shared__ __align__(8) uint64_t s_full[16], s_empty[16];
    __shared__ float s_red[MT][GB_CONSUMERS];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_units = SWIGLU ? p.N / 2 : p.N;
    const int ipu = (SWIGLU ? 2 : 1) * p.nseg;                 // items per unit
    const int TW = gridDim.x * GB_CONSUMERS;
    auto units_of = [&](int gl) { return gl < n_units ? (n_units - gl + TW - 1) / TW : 0; };
    // rows of a unit: SWIGLU -> (gate, up) of output column `unit`; else the row itself
    auto row_of = [&](int unit, int sel) {
        if (SWIGLU) return (unit / 128) * 256 + (unit % 128) + sel * 128;
        return unit;
    };

    if (tid == 0) {
        for (int s = 0; s < p.stages; ++s) {
            ptx::mbar_init(ptx::smem_u32(&s_full[s]), 1);
            ptx::mbar_init(ptx::smem_u32(&s_empty[s]), GB_CONSUMERS);
        }
        ptx::fence_mbar_init();
    }
    __syncthreads();

    if (warp == GB_CONSUMERS) {
        // ------------------------------------------------------------ producer: weights do not depend on the
        // previous kernel, so the stream starts before griddepcontrol.wait (PDL) and before x is staged.
        if (lane == 0) {
            int max_items = 0;
            for (int w = 0; w < GB_CONSUMERS; ++w) {
                const int it = units_of(blockIdx.x * GB_CONSUMERS + w) * ipu;
                max_items = it > max_items ? it : max_items;
            }
            int stage = 0;
            uint32_t phase = 0;
            for (int i = 0; i < max_items; ++i) {
                ptx::mbar_wait(ptx::smem_u32(&s_empty[stage]), phase ^ 1);
                const uint32_t bar = ptx::smem_u32(&s_full[stage]);
                uint32_t tx = 0;
                for (int w = 0; w < GB_CONSUMERS; ++w)
                    if (i < units_of(blockIdx.x * GB_CONSUMERS + w) * ipu) tx += seg_bytes;
                ptx::mbar_arrive_expect_tx(bar, tx);
                for (int w = 0; w < GB_CONSUMERS; ++w) {
                    const int gl = blockIdx.x * GB_CONSUMERS + w;
                    if (i >= units_of(gl) * ipu) continue;
                    const int unit = gl + (i / ipu) * TW;
                    const int r = i % ipu;
                    const int row = row_of(unit, r / p.nseg);
                    const int seg = r % p.nseg;
                    bulk_g2s(ring_u32 + stage * stage_bytes + w * seg_bytes, p.W + (size_t)row * p.ldw + (size_t)seg * p.seg_len,
                             seg_bytes, bar);
                }
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ------------------------------------------------------------ consumers
        pdl_launch_dependents();
        pdl_wait();
        const int ctid = tid;                                   // 0..255
        const int kv = K / 8;
        float ss[MT];
#pragma unroll
        for (int m = 0; m < MT; ++m) ss[m] = 0.f;
        for (int i = ctid; i < kv; i += 32 * GB_CONSUMERS) {
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                uint4 v = *(reinterpret_cast<const uint4*>(p.x + (size_t)m * p.ldx) + i);
                reinterpret_cast<uint4*>(sx + (size_t)m * K)[i] = v;
                if (p.norm_w != nullptr) {
                    float2 f;
                    f = unpack_bf16(v.x); ss[m] += f.x * f.x + f.y * f.y;
                    f = unpack_bf16(v.y); ss[m] += f.x * f.x + f.y * f.y;
                    f = unpack_bf16(v.z); ss[m] += f.x * f.x + f.y * f.y;
                    f = unpack_bf16(v.w); ss[m] += f.x * f.x + f.y * f.y;
                }
            }
        }
        if (p.norm_w != nullptr) {
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                float v = wsum_b(ss[m]);
                if (lane == 0) s_red[m][warp] = v;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");       // consumers only
            float rstd[MT];
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                float t = 0.f;
#pragma unroll
                for (int w = 0; w < GB_CONSUMERS; ++w) t += s_red[m][w];
                rstd[m] = rsqrtf(t / K + p.eps);
            }
            for (int i = ctid; i < kv; i += 32 * GB_CONSUMERS) {
                uint4 wv = __ldg(reinterpret_cast<const uint4*>(p.norm_w) + i);
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
        asm volatile("bar.sync 1, 256;" ::: "memory");

        const int gl = blockIdx.x * GB_CONSUMERS + warp;
        const int my_items = units_of(gl) * ipu;
        int max_items = 0;
        for (int w = 0; w < GB_CONSUMERS; ++w) {
            const int it = units_of(blockIdx.x * GB_CONSUMERS + w) * ipu;
            max_items = it > max_items ? it : max_items;
        }
        const int chunks = p.seg_len / 8;                        // uint4 per segment
        float acc[2][MT];
#pragma unroll
        for (int m = 0; m < MT; ++m) { acc[0][m] = 0.f; acc[1][m] = 0.f; }
        int stage = 0;
        uint32_t phase = 0;
        for (int i = 0; i < max_items; ++i) {
            ptx::mbar_wait(ptx::smem_u32(&s_full[stage]), phase);
            if (i < my_items) {
                const int r = i % ipu;
                const int sel = r / p.nseg, seg = r % p.nseg;
                const uint4* wseg = reinterpret_cast<const uint4*>(ring + (size_t)stage * stage_bytes + (size_t)warp * seg_bytes);
                float part[MT];
#pragma unroll
                for (int m = 0; m < MT; ++m) part[m] = 0.f;
#pragma unroll 4
                for (int c = lane; c < chunks; c += 32) {
                    const uint4 wv = wseg[c];
#pragma unroll
                    for (int m = 0; m < MT; ++m)
                        part[m] += dot8b(wv, reinterpret_cast<const uint4*>(sx + (size_t)m * K + (size_t)seg * p.seg_len)[c]);
                }
#pragma unroll
                for (int m = 0; m < MT; ++m) {
                    if (sel == 0) acc[0][m] += part[m]; else acc[1][m] += part[m];
                }
            }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&s_empty[stage]));   // this warp is done with its slot
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
            if (i < my_items && (i % ipu) == ipu - 1) {
                // ---- unit finished
                const int unit = gl + (i / ipu) * TW;
#pragma unroll
                for (int m = 0; m < MT; ++m) { acc[0][m] = wsum_b(acc[0][m]); if (SWIGLU) acc[1][m] = wsum_b(acc[1][m]); }
                if (lane == 0) {
#pragma unroll
                    for (int m = 0; m < MT; ++m) {
                        if (SWIGLU) {
                            const float g = bf16r(acc[0][m]), u = bf16r(acc[1][m]);
                            reinterpret_cast<__nv_bfloat16*>(p.out)[(size_t)m * p.ldo + unit] = __float2bfloat16_rn(u * bf16r(silu_f(g)));
                        } else {
                            float y = acc[0][m];
                            if (p.bias) y += __bfloat162float(p.bias[unit]);
                            y = bf16r(y);
                            if (p.residual) y = bf16r(y + __bfloat162float(p.residual[(size_t)m * p.ldr + unit]));
                            if (p.out_f32) reinterpret_cast<float*>(p.out)[(size_t)m * p.ldo + unit] = y;
                            else reinterpret_cast<__nv_bfloat16*>(p.out)[(size_t)m * p.ldo + unit] = __float2bfloat16_rn(y);
                        }
                    }
                }
#pragma unroll
                for (int m = 0; m < MT; ++m) { acc[0][m] = 0.f; acc[1][m] = 0.f; }
            }
        }
    }
}

}  // namespace

int gemv_bulk_bf16(const __nv_bfloat16* x, int ldx, const __nv_bfloat16* W, int ldw, void* out, int ldo, int M, int N,
                   int K, const __nv_bfloat16* norm_w, float eps, const __nv_bfloat16* bias,
                   const __nv_bfloat16* residual, int ldr, int act, int out_f32, cudaStream_t s) {
    if (M < 1 || M > GB_MAXM || K % 256 != 0 || (act != 0 && act != 3)) return GVL_ERR_ARG;
    if (act == 3 && N % 256 != 0) return GVL_ERR_ARG;
    if ((ldw * 2) % 16 != 0 || (reinterpret_cast<uintptr_t>(W) & 15)) return GVL_ERR_ALIGN;
    GbParams p;
    p.x = x; p.ldx = ldx; p.W = W; p.ldw = ldw; p.out = out; p.ldo = ldo; p.N = N; p.K = K;
    p.norm_w = norm_w; p.eps = eps; p.bias = bias; p.residual = residual; p.ldr = ldr; p.out_f32 = out_f32;
    p.nseg = (K + GB_SEG_MAX - 1) / GB_SEG_MAX;
    while (K % (p.nseg * 256) != 0) ++p.nseg;           // segments must be whole 256-element lane strides
    p.seg_len = K / p.nseg;
    const size_t x_bytes = ((size_t)M * K * 2 + 127) & ~size_t(127);
    const size_t stage_bytes = (size_t)GB_CONSUMERS * p.seg_len * 2;
    int stages = (int)((GB_SMEM_BUDGET - x_bytes) / stage_bytes);
    if (stages > 16) stages = 16;
    if (stages < 2) return GVL_ERR_ARG;
    p.stages = stages;
    const size_t smem = x_bytes + (size_t)stages * stage_bytes;
    prof_begin(GVL_PROF_GEMV, 2.0 * (double)N * K, s);
    const int units = act == 3 ? N / 2 : N;
    int grid = (units + GB_CONSUMERS - 1) / GB_CONSUMERS;
    if (grid > num_sms()) grid = num_sms();
#define GB_LAUNCH(MT, SW)                                                                                           \
    do {                                                                                                            \
        auto kern = gemv_bulk_kernel<MT, SW>;                                                                       \
        static size_t max_set = 0;                                                                                  \
        if (smem > max_set) {                                                                                       \
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) \
                return GVL_ERR_CUDA;                                                                                \
            max_set = smem;                                                                                         \
        }                                                                                                           \
        if (launch_k(kern, dim3(grid), dim3(GB_THREADS), smem, s, p) != cudaSuccess) return GVL_ERR_CUDA;           \
    } while (0)
    if (act == 3) {
        switch (M) { case 1: GB_LAUNCH(1, true); break; case 2: GB_LAUNCH(2, true); break;
                     case 3: GB_LAUNCH(3, true); break; default: GB_LAUNCH(4, true); break; }
    } else {
        switch (M) { case 1: GB_LAUNCH(1, false); break; case 2: GB_LAUNCH(2, false); break;
                     case 3: GB_LAUNCH(3, false); break; default: GB_LAUNCH(4, false); break; }
    }
#undef GB_LAUNCH
    prof_end(GVL_PROF_GEMV, s);
    g_launch_count++;
    return cudaGetLastError() == cudaSuccess ? GVL_OK : GVL_ERR_CUDA;
}

}  // namespace gvl
