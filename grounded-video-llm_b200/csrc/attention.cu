// Fused softmax(Q K^T * scale) V for the three prefill-side attention shapes on the hot path:
//   CLIP      non-causal, d=64,  S=577   (eager bmm/softmax/bmm, modeling_clip.py:252-328)
//   IV2       non-causal, d=88,  S=2049  (flash_attn_varlen_qkvpacked, internvideo2.py:493-538, 585-605)
//   Phi/Llama causal,     d=96/128       (flash_attn_func, modeling_phi3.py:778-876; eager twin :531-610)
// Scores never touch HBM (the reference's eager CLIP path materialises 12*16*577^2 of them).
//
// v1 data path: cp.async double-buffered K/V tiles -> ldmatrix -> mma.sync.m16n8k16 (bf16, fp32
// accumulate) with FlashAttention-2 style online softmax in registers. One CTA = 128 query rows
// (8 warps x 16 rows) of one (batch, head); KV tiles of 64 tokens.
// Numerics follow the reference: fp32 scores, fp32 softmax statistics, probabilities rounded to
// bf16 before the P*V product, fp32 output accumulation, one final bf16 rounding.
#include "gvl_internal.h"
#include <stdlib.h>
#include "ptx.cuh"

namespace gvl {

namespace {

constexpr int ATT_BM = 128;
constexpr int ATT_BN = 64;
constexpr int ATT_THREADS = 256;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool pred) {
    int sz = pred ? 16 : 0;  // src-size 0 => zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// HD = padded head dim (multiple of 16); rows in smem have stride HD+8 elements so that the 8 rows
// an ldmatrix touches fall in 8 distinct 16-byte bank groups.
template <int HD, bool CAUSAL>
__global__ void __launch_bounds__(ATT_THREADS)
attn_fwd_kernel(const AttnArgs a) {
    constexpr int LDS = HD + 8;
    constexpr int CH = HD / 8;  // 16-byte chunks per row
    extern __shared__ __align__(16) uint8_t smem[];
    __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem);       // [128][LDS]
    __nv_bfloat16* sK = sQ + ATT_BM * LDS;                             // [2][64][LDS]
    __nv_bfloat16* sV = sK + 2 * ATT_BN * LDS;                         // [2][64][LDS]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int m0 = blockIdx.x * ATT_BM;
    const int h = blockIdx.y, b = blockIdx.z;
    const int hk = h / (a.heads / a.kv_heads);
    const int d_real = a.head_dim;
    const int real_ch = d_real / 8;

    const __nv_bfloat16* qp = a.q + b * a.q_bs + h * a.q_hs;
    const __nv_bfloat16* kp = a.k + b * a.k_bs + hk * a.k_hs;
    const __nv_bfloat16* vp = a.v + b * a.v_bs + hk * a.v_hs;

    const int causal_off = a.skv - a.sq;
    int kv_end = a.skv;
    if (CAUSAL) {
        int last = m0 + ATT_BM - 1 + causal_off + 1;
        kv_end = last < a.skv ? last : a.skv;
        if (kv_end < 0) kv_end = 0;
    }
    const int n_blocks = (kv_end + ATT_BN - 1) / ATT_BN;

    // ---- async load Q tile
    for (int i = tid; i < ATT_BM * CH; i += ATT_THREADS) {
        int r = i / CH, c = i % CH;
        bool ok = (m0 + r) < a.sq && c < real_ch;
        const __nv_bfloat16* src = qp + (long long)(ok ? (m0 + r) : 0) * a.q_ts + (ok ? c * 8 : 0);
        cp_async16(ptx::smem_u32(sQ + r * LDS + c * 8), src, ok);
    }
    auto load_kv = [&](int blk, int buf) {
        const int n0 = blk * ATT_BN;
        for (int i = tid; i < ATT_BN * CH; i += ATT_THREADS) {
            int r = i / CH, c = i % CH;
            bool ok = (n0 + r) < a.skv && c < real_ch;
            long long tok = ok ? (n0 + r) : 0;
            int co = ok ? c * 8 : 0;
            cp_async16(ptx::smem_u32(sK + (buf * ATT_BN + r) * LDS + c * 8), kp + tok * a.k_ts + co, ok);
            cp_async16(ptx::smem_u32(sV + (buf * ATT_BN + r) * LDS + c * 8), vp + tok * a.v_ts + co, ok);
        }
    };
    if (n_blocks > 0) load_kv(0, 0);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    // ---- Q fragments to registers (A operand, 16 rows per warp)
    uint32_t qf[HD / 16][4];
    {
        const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int cbase = (lane >> 4) * 8;
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk)
            ldsm_x4(ptx::smem_u32(sQ + r * LDS + kk * 16 + cbase), qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3]);
    }

    float o[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
    float m_run[2] = {-INFINITY, -INFINITY};
    float l_run[2] = {0.f, 0.f};
    const float sl2 = a.scale * 1.4426950408889634f;
    const int qrow0 = m0 + warp * 16 + g;  // rows qrow0 and qrow0 + 8

    for (int blk = 0; blk < n_blocks; ++blk) {
        const int buf = blk & 1;
        if (blk + 1 < n_blocks) load_kv(blk + 1, buf ^ 1);
        cp_async_commit();

        // ---- S = Q K^T for this KV tile: 16 x 64 per warp
        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
        const __nv_bfloat16* kb = sK + buf * ATT_BN * LDS;
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
#pragma unroll
            for (int np = 0; np < 4; ++np) {  // pairs of 8-wide n blocks
                uint32_t b0, b1, b2, b3;
                const int n = np * 16 + (lane & 7) + (lane >> 4) * 8;
                const int kc = kk * 16 + ((lane >> 3) & 1) * 8;
                ldsm_x4(ptx::smem_u32(kb + n * LDS + kc), b0, b1, b2, b3);
                mma16816(s[np * 2], qf[kk], b0, b1);
                mma16816(s[np * 2 + 1], qf[kk], b2, b3);
            }
        }

        // ---- mask + online softmax (log2 domain)
        const int n0 = blk * ATT_BN;
        const bool need_mask = (n0 + ATT_BN > a.skv) || (CAUSAL && (n0 + ATT_BN - 1 > m0 + warp * 16 + causal_off));
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float x = s[i][e];
                if (a.round_scores) x = bf16r(x);
                x *= sl2;
                if (need_mask) {
                    const int col = n0 + i * 8 + t4 * 2 + (e & 1);
                    const int row = qrow0 + (e >> 1) * 8;
                    if (col >= a.skv || (CAUSAL && col > row + causal_off)) x = -INFINITY;
                }
                s[i][e] = x;
                mx[e >> 1] = fmaxf(mx[e >> 1], x);
            }
        }
        float corr[2], msub[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            const float m_new = fmaxf(m_run[r], mx[r]);
            msub[r] = (m_new == -INFINITY) ? 0.f : m_new;
            corr[r] = exp2f(m_run[r] - msub[r]);
            m_run[r] = m_new;
        }
        float rs[2] = {0.f, 0.f};
        uint32_t pf[4][4];  // P as A-operand fragments: 4 k-steps of 16 kv tokens
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float p0 = exp2f(s[i][0] - msub[0]);
            float p1 = exp2f(s[i][1] - msub[0]);
            float p2 = exp2f(s[i][2] - msub[1]);
            float p3 = exp2f(s[i][3] - msub[1]);
            // the row sum uses the bf16-rounded probabilities that feed the P*V product
            p0 = bf16r(p0); p1 = bf16r(p1); p2 = bf16r(p2); p3 = bf16r(p3);
            rs[0] += p0 + p1;
            rs[1] += p2 + p3;
            const int ks = i >> 1;
            if ((i & 1) == 0) {
                pf[ks][0] = pack_bf16(p0, p1);
                pf[ks][1] = pack_bf16(p2, p3);
            } else {
                pf[ks][2] = pack_bf16(p0, p1);
                pf[ks][3] = pack_bf16(p2, p3);
            }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * corr[r] + rs[r];
#pragma unroll
        for (int i = 0; i < HD / 8; ++i) {
            o[i][0] *= corr[0]; o[i][1] *= corr[0];
            o[i][2] *= corr[1]; o[i][3] *= corr[1];
        }

        // ---- O += P V
        const __nv_bfloat16* vb = sV + buf * ATT_BN * LDS;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
            for (int dp = 0; dp < HD / 16; ++dp) {
                uint32_t b0, b1, b2, b3;
                const int kv = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int dc = dp * 16 + (lane >> 4) * 8;
                ldsm_x4_t(ptx::smem_u32(vb + kv * LDS + dc), b0, b1, b2, b3);
                mma16816(o[dp * 2], pf[ks], b0, b1);
                mma16816(o[dp * 2 + 1], pf[ks], b2, b3);
            }
        }
        cp_async_wait<0>();
        __syncthreads();
    }

    // ---- finalize: O / l, stage through smem (Q tile is dead) for 16-byte coalesced stores
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
    const float inv0 = l_run[0] > 0.f ? 1.f / l_run[0] : 0.f;
    const float inv1 = l_run[1] > 0.f ? 1.f / l_run[1] : 0.f;
    {
        const int r = warp * 16 + g;
#pragma unroll
        for (int i = 0; i < HD / 8; ++i) {
            const int c = i * 8 + t4 * 2;
            *reinterpret_cast<uint32_t*>(sQ + r * LDS + c) = pack_bf16(o[i][0] * inv0, o[i][1] * inv0);
            *reinterpret_cast<uint32_t*>(sQ + (r + 8) * LDS + c) = pack_bf16(o[i][2] * inv1, o[i][3] * inv1);
        }
    }
    __syncwarp();
    __nv_bfloat16* op = a.o + b * a.o_bs + h * a.o_hs;
    const int out_ch = (a.o_dim > 0 ? a.o_dim : d_real) / 8;
    for (int i = lane; i < 16 * out_ch; i += 32) {
        const int r = warp * 16 + i / out_ch, c = i % out_ch;
        if (m0 + r < a.sq) {
            uint4 val = *reinterpret_cast<const uint4*>(sQ + r * LDS + c * 8);
            *reinterpret_cast<uint4*>(op + (long long)(m0 + r) * a.o_ts + c * 8) = val;
        }
    }
}

template <int HD, bool CAUSAL>
int launch_attn(const AttnArgs& a, cudaStream_t stream) {
    constexpr int LDS = HD + 8;
    constexpr int SMEM = (ATT_BM + 4 * ATT_BN) * LDS * 2;
    auto kern = attn_fwd_kernel<HD, CAUSAL>;
    static unsigned long long attr_devs = 0ull;
    if (first_use_on_device(attr_devs)) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess)
            return GVL_ERR_CUDA;
    }
    dim3 grid((a.sq + ATT_BM - 1) / ATT_BM, a.heads, a.batch);
    kern<<<grid, ATT_THREADS, SMEM, stream>>>(a);
    g_launch_count++;
    return cudaGetLastError() == cudaSuccess ? GVL_OK : GVL_ERR_CUDA;
}

}  // namespace

int attention_fwd(const AttnArgs& a, cudaStream_t stream) {
    if (a.head_dim % 8 != 0 || a.head_dim > 128 || a.heads % a.kv_heads != 0) return GVL_ERR_ARG;
    if (a.sq <= 0 || a.skv <= 0) return GVL_ERR_ARG;
    if (attention_tc_supported(a)) {
        prof_begin(GVL_PROF_ATTN, 4.0 * a.batch * a.heads * (double)a.sq * a.skv * a.head_dim * (a.causal ? 0.5 : 1.0), stream);
        // two query tiles per CTA (attention_tc2.cu) whenever there is more than one tile of queries; the one-tile kernel
        // (attention_tc.cu) serves Sq <= 128 (decode-time prefill of short prompts, tests)
        int rc = a.sq > 128 ? attention_tc2_fwd(a, stream) : attention_tc_fwd(a, stream);
        prof_end(GVL_PROF_ATTN, stream);
        return rc;
    }
    return attention_mma_fwd(a, stream);
}

int attention_mma_fwd(const AttnArgs& a, cudaStream_t stream) {
    const int hd = (a.head_dim + 15) / 16 * 16;
    // algorithmic FLOPs: 4*Sq*Skv*D per (batch, head), halved under the causal mask
    prof_begin(GVL_PROF_ATTN, 4.0 * a.batch * a.heads * (double)a.sq * a.skv * a.head_dim * (a.causal ? 0.5 : 1.0), stream);
    int rc;
    if (a.causal) {
        if (hd <= 64) rc = launch_attn<64, true>(a, stream);
        else if (hd <= 96) rc = launch_attn<96, true>(a, stream);
        else rc = launch_attn<128, true>(a, stream);
    } else {
        if (hd <= 64) rc = launch_attn<64, false>(a, stream);
        else if (hd <= 96) rc = launch_attn<96, false>(a, stream);
        else rc = launch_attn<128, false>(a, stream);
    }
    prof_end(GVL_PROF_ATTN, stream);
    return rc;
}

}  // namespace gvl
