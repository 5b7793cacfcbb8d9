// Fused GEMM epilogue on a 32-column chunk of one accumulator row (registers after tcgen05.ld), shared by the
// 1-CTA and the 2-CTA (cta_group::2) tcgen05 GEMM kernels. Rounding points mirror the reference's bf16 autocast
// forward (SURVEY 8a "numerics contract"): the fp32 accumulator (+bias) is rounded to bf16 first (what nn.Linear
// returns), then activation / LayerScale / residual are applied with the reference's intermediate roundings.
#pragma once
#include "gvl_internal.h"
#include "ptx.cuh"

namespace gvl {

struct GemmParams2 {
    int M, N, K;
    const __nv_bfloat16* bias;  // [N] or nullptr
    const float* gamma;         // [N_out] LayerScale or nullptr
    const void* residual;       // [M, ldr] (bf16 if RES==1, f32 if RES==2); may alias out
    void* out;                  // [M, ldo] bf16 or f32
    int ldo, ldr;
    int num_m_tiles, num_n_tiles;
};

enum { EPI_ACT_NONE = 0, EPI_ACT_GELU = 1, EPI_ACT_QUICKGELU = 2, EPI_ACT_SWIGLU = 3 };
enum { EPI_RES_NONE = 0, EPI_RES_BF16 = 1, EPI_RES_F32 = 2 };

// The 32 bias values of a chunk (4 x 16 bytes; zeros where there is no bias / past N). Split from the math so that a caller can
// issue these loads a chunk ahead: with 2 epilogue warps per scheduler nothing else hides their latency (the IV2 fc1 capture had
// the epilogue warps on stall_long_sb at the first bias use, profiles/r2_gemm.md).
__device__ __forceinline__ void epilogue_load_bias(const GemmParams2& p, int col_in, uint4 (&b4)[4]) {
#pragma unroll
    for (int g8 = 0; g8 < 4; ++g8) b4[g8] = make_uint4(0u, 0u, 0u, 0u);
    if (p.bias != nullptr && col_in < p.N) {
        const uint4* bp = reinterpret_cast<const uint4*>(p.bias + col_in);
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8)
            if (col_in + g8 * 8 < p.N) b4[g8] = __ldg(bp + g8);
    }
}

// The 32 bf16 residual values of a chunk (RES_BF16), loadable a chunk ahead for the same reason. The residual may alias the
// output (in-place residual stream): a chunk is read before ITS columns are written, other chunks touch other columns.
__device__ __forceinline__ void epilogue_load_res_bf16(const GemmParams2& p, int row, bool row_ok, int col_out, int n_out_total,
                                                       uint4 (&r4)[4]) {
#pragma unroll
    for (int g8 = 0; g8 < 4; ++g8) r4[g8] = make_uint4(0u, 0u, 0u, 0u);
    if (row_ok && col_out < n_out_total) {
        const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.residual) + size_t(row) * p.ldr + col_out);
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8)
            if (col_out + g8 * 8 < n_out_total) r4[g8] = rp[g8];
    }
}

// acc: 32 fp32 accumulator columns starting at GEMM column col_in (SWIGLU: the gate columns; accu = the matching up
// columns). Writes 32 output columns starting at col_out of row `row`. b4: epilogue_load_bias(p, col_in);
// r4: epilogue_load_res_bf16(...) when RES == EPI_RES_BF16 (ignored otherwise).
template <int ACT, int RES, bool OUT_F32>
__device__ __forceinline__ void epilogue_chunk32(const GemmParams2& p, const uint32_t (&acc)[32], const uint32_t (&accu)[32],
                                                 int row, bool row_ok, int col_in, int col_out, int n_out_total,
                                                 const uint4 (&b4)[4], const uint4 (&r4)[4]) {
    float v[32];
    if (ACT == EPI_ACT_SWIGLU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            // Phi3MLP (modeling_phi3.py:458-464): up * silu(gate) on bf16 tensors
            const float g = bf16r(__uint_as_float(acc[j]));
            const float u = bf16r(__uint_as_float(accu[j]));
            v[j] = bf16r(u * bf16r(silu_f(g)));
        }
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
        if (p.bias != nullptr) {                       // x + 0.0f is exact, so the zero-filled groups need no guard
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) {
                const uint4 b = b4[g8];
                float2 f;
                f = unpack_bf16(b.x); v[g8 * 8 + 0] += f.x; v[g8 * 8 + 1] += f.y;
                f = unpack_bf16(b.y); v[g8 * 8 + 2] += f.x; v[g8 * 8 + 3] += f.y;
                f = unpack_bf16(b.z); v[g8 * 8 + 4] += f.x; v[g8 * 8 + 5] += f.y;
                f = unpack_bf16(b.w); v[g8 * 8 + 6] += f.x; v[g8 * 8 + 7] += f.y;
            }
        }
        // the activation result is a bf16 tensor in the reference; when nothing but the final pack follows (no LayerScale, no
        // residual) that pack IS the rounding, so the explicit round-trip is skipped (bf16r is idempotent)
        const bool round_act = (p.gamma != nullptr) || RES != EPI_RES_NONE || OUT_F32;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            float x = bf16r(v[j]);  // what nn.Linear returns under bf16 autocast
            if (ACT == EPI_ACT_GELU) x = gelu_erf(x);
            if (ACT == EPI_ACT_QUICKGELU) x = quick_gelu_bf16(x);
            v[j] = x;
        }
        if (ACT == EPI_ACT_GELU && round_act) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = bf16r(v[j]);
        }
    }
    if (!row_ok || col_out >= n_out_total) return;
    if (p.gamma != nullptr) {
        // LayerScale (internvideo2.py:451-466): fp32 multiply, rounded back to bf16
#pragma unroll
        for (int g4 = 0; g4 < 8; ++g4) {
            if (col_out + g4 * 4 < n_out_total) {
                const float4 gm = __ldg(reinterpret_cast<const float4*>(p.gamma + col_out) + g4);
                v[g4 * 4 + 0] = bf16r(v[g4 * 4 + 0] * gm.x);
                v[g4 * 4 + 1] = bf16r(v[g4 * 4 + 1] * gm.y);
                v[g4 * 4 + 2] = bf16r(v[g4 * 4 + 2] * gm.z);
                v[g4 * 4 + 3] = bf16r(v[g4 * 4 + 3] * gm.w);
            }
        }
    }
    if (RES == EPI_RES_BF16) {
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
            const uint4 r = r4[g8];                    // zero-filled past the edge (those columns are not stored)
            float2 f;
            f = unpack_bf16(r.x); v[g8 * 8 + 0] += f.x; v[g8 * 8 + 1] += f.y;
            f = unpack_bf16(r.y); v[g8 * 8 + 2] += f.x; v[g8 * 8 + 3] += f.y;
            f = unpack_bf16(r.z); v[g8 * 8 + 4] += f.x; v[g8 * 8 + 5] += f.y;
            f = unpack_bf16(r.w); v[g8 * 8 + 6] += f.x; v[g8 * 8 + 7] += f.y;
        }
    } else if (RES == EPI_RES_F32) {
        const float4* rp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.residual) +
                                                           size_t(row) * p.ldr + col_out);
#pragma unroll
        for (int g4 = 0; g4 < 8; ++g4) {
            if (col_out + g4 * 4 < n_out_total) {
                const float4 r = rp[g4];
                v[g4 * 4 + 0] += r.x; v[g4 * 4 + 1] += r.y;
                v[g4 * 4 + 2] += r.z; v[g4 * 4 + 3] += r.w;
            }
        }
    }
    if (OUT_F32) {
        float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + size_t(row) * p.ldo + col_out);
#pragma unroll
        for (int g4 = 0; g4 < 8; ++g4)
            if (col_out + g4 * 4 < n_out_total) op[g4] = make_float4(v[g4 * 4], v[g4 * 4 + 1], v[g4 * 4 + 2], v[g4 * 4 + 3]);
    } else {
        uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + size_t(row) * p.ldo + col_out);
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
            if (col_out + g8 * 8 < n_out_total) {
                uint4 o;
                o.x = pack_bf16(v[g8 * 8 + 0], v[g8 * 8 + 1]);
                o.y = pack_bf16(v[g8 * 8 + 2], v[g8 * 8 + 3]);
                o.z = pack_bf16(v[g8 * 8 + 4], v[g8 * 8 + 5]);
                o.w = pack_bf16(v[g8 * 8 + 6], v[g8 * 8 + 7]);
                op[g8] = o;
            }
        }
    }
}

// convenience form: bias loaded in place (1-CTA kernel, SwiGLU path)
template <int ACT, int RES, bool OUT_F32>
__device__ __forceinline__ void epilogue_chunk32(const GemmParams2& p, const uint32_t (&acc)[32], const uint32_t (&accu)[32],
                                                 int row, bool row_ok, int col_in, int col_out, int n_out_total) {
    uint4 b4[4], r4[4];
    epilogue_load_bias(p, col_in, b4);
    if (RES == EPI_RES_BF16) epilogue_load_res_bf16(p, row, row_ok, col_out, n_out_total, r4);
    epilogue_chunk32<ACT, RES, OUT_F32>(p, acc, accu, row, row_ok, col_in, col_out, n_out_total, b4, r4);
}

}  // namespace gvl
