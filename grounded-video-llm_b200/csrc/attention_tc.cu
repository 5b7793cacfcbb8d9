// tcgen05 / TMEM / TMA fused attention forward for head dims 64 / 96 / 128 (prefill-side shapes of the path):
//   CLIP   non-causal d=64  S=577  (modeling_clip.py:252-328)
//   IV2    non-causal d=88 stored padded to 96 by the qkv GEMM (internvideo2.py:493-538, 585-605)
//   Phi-3  causal d=96, Llama-3 causal d=128 GQA (modeling_phi3.py:778-876, modeling_llama.py:537-594)
//
// One CTA = 128 query rows of one (batch, head). Warp roles:
//   warp 0      TMA producer: Q once, then K_j / V_j tiles (128 kv rows) into a multi-stage smem ring.
//               Every operand is stored as head_dim/32 chunks of [rows x 64 B] with TMA SWIZZLE_64B.
//   warp 1      MMA issuer (one thread): S_j = Q K_j^T  (SS, both K-major, M=128 N=128 K=hd) into TMEM,
//               O += P_j V_j (TS: A = P_j bf16 in TMEM, B = V_j MN-major from smem, N=hd).
//               Issue order QK_0, QK_1, PV_0, QK_2, PV_1, ... so the tensor pipe works on QK_{j+1} / PV_{j-1}
//               while the softmax warps work on S_j.
//   warps 2..5  softmax: one query row per thread (TMEM lane), two passes over S_j in TMEM
//               (row max; exp2 + row sum + bf16 P written back over S_j), lazy O rescaling
//               (only when the running max grows by more than 2^8), final O / l epilogue.
// TMEM columns: S0 [0,128) S1 [128,256) O [256,256+hd); P_j aliases the first 64 columns of S_j.
//
// Numerics as in attention.cu / the reference flash-attn path: fp32 scores, fp32 statistics, P rounded to bf16
// before P@V, fp32 accumulation, one final rounding. The row sum uses the un-rounded fp32 probabilities (FA2).
#include <cuda.h>
#include <type_traits>
#include "gvl_internal.h"
#include "ptx.cuh"

namespace gvl {

namespace {

constexpr int TQ = 128;   // query rows per CTA
constexpr int TK = 128;   // kv rows per tile
constexpr int ATC_THREADS = 192;
constexpr float RESCALE_LOG2 = 8.0f;

template <int HD>
struct AtcCfg {
    static constexpr int CH = HD / 32;                  // 64-byte chunks per row
    static constexpr int Q_BYTES = TQ * HD * 2;
    static constexpr int KV_BYTES = TK * HD * 2;        // one K or one V tile
    static constexpr int STAGES = (HD <= 96) ? 3 : 2;
    static constexpr int SMEM = Q_BYTES + STAGES * 2 * KV_BYTES + 1024 + 256;
    static constexpr int CHUNK_BYTES = 128 * 64;        // [128 rows x 64 B]
};

// K-major operand chunk [rows x 64 B], SWIZZLE_64B: 8-row groups are 512 B apart.
__device__ __forceinline__ uint64_t desc_kmajor_sw64(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;                 // LBO unused for swizzled K-major
    d |= (uint64_t)(512 >> 4) << 32;        // SBO
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;                 // SWIZZLE_64B
    return d;
}
// MN-major operand (V: rows = kv (K index), 64-byte rows = 32 contiguous d (MN index)), SWIZZLE_64B:
// LBO = stride between 32-element MN chunks (one chunk = 128 rows x 64 B), SBO = stride between 8-row K groups.
__device__ __forceinline__ uint64_t desc_mnmajor_sw64(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)(8192 >> 4) << 16;       // LBO = chunk stride
    d |= (uint64_t)(512 >> 4) << 32;        // SBO = 8 rows x 64 B
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(b_mn_major) << 16) | (uint32_t(N >> 3) << 17) |
           (uint32_t(M >> 4) << 24);
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

struct AtcParams {
    __nv_bfloat16* o;
    long long o_bs, o_ts, o_hs;
    int sq, skv, heads, kv_heads, o_dim;
    float scale_log2;
    int round_scores;
};

template <int HD, bool CAUSAL, bool ROUND>
__global__ void __launch_bounds__(ATC_THREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const AtcParams p) {
    using Cfg = AtcCfg<HD>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int CH = Cfg::CH;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sQ = base;
    const uint32_t sKV = base + Cfg::Q_BYTES;     // stage s: K at sKV + s*2*KV_BYTES, V right after
    const uint32_t bars = sKV + STAGES * 2 * Cfg::KV_BYTES;
    // barrier map
    const uint32_t q_full = bars;
    auto k_full = [&](int s) { return bars + 8u * (1 + s); };
    auto v_full = [&](int s) { return bars + 8u * (1 + STAGES + s); };
    auto k_empty = [&](int s) { return bars + 8u * (1 + 2 * STAGES + s); };
    auto v_empty = [&](int s) { return bars + 8u * (1 + 3 * STAGES + s); };
    auto s_full = [&](int b) { return bars + 8u * (1 + 4 * STAGES + b); };
    auto p_ready = [&](int b) { return bars + 8u * (3 + 4 * STAGES + b); };
    const uint32_t o_done = bars + 8u * (5 + 4 * STAGES);
    const uint32_t tmem_slot = bars + 8u * (6 + 4 * STAGES);
    uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * TQ;
    const int h = blockIdx.y, b = blockIdx.z;
    const int hk = h / (p.heads / p.kv_heads);
    const int causal_off = p.skv - p.sq;
    int kv_end = p.skv;
    if (CAUSAL) {
        const int last = m0 + TQ + causal_off;   // exclusive bound for the last row of the tile
        kv_end = last < p.skv ? last : p.skv;
        if (kv_end < 1) kv_end = 1;
    }
    const int n_tiles = (kv_end + TK - 1) / TK;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmQ);
        ptx::prefetch_tmap(&tmK);
        ptx::prefetch_tmap(&tmV);
        ptx::mbar_init(q_full, 1);
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(k_full(s), 1);
            ptx::mbar_init(v_full(s), 1);
            ptx::mbar_init(k_empty(s), 1);
            ptx::mbar_init(v_empty(s), 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(s_full(i), 1);
            ptx::mbar_init(p_ready(i), 4);
        }
        ptx::mbar_init(o_done, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, 512);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *tmem_slot_ptr;
    const uint32_t tS0 = tmem, tO = tmem + 256;

    if (warp == 0) {
        if (lane == 0) {
            // ---------------------------------------------------------------- TMA producer
            ptx::mbar_arrive_expect_tx(q_full, Cfg::Q_BYTES);
            for (int c = 0; c < CH; ++c) tma_load_4d(sQ + c * Cfg::CHUNK_BYTES, &tmQ, q_full, c * 32, m0, h, b);
            int stage = 0;
            uint32_t phase = 0;
            for (int j = 0; j < n_tiles; ++j) {
                const uint32_t sk = sKV + stage * 2 * Cfg::KV_BYTES;
                const uint32_t sv = sk + Cfg::KV_BYTES;
                ptx::mbar_wait(k_empty(stage), phase ^ 1);
                ptx::mbar_arrive_expect_tx(k_full(stage), Cfg::KV_BYTES);
                for (int c = 0; c < CH; ++c) tma_load_4d(sk + c * Cfg::CHUNK_BYTES, &tmK, k_full(stage), c * 32, j * TK, hk, b);
                ptx::mbar_wait(v_empty(stage), phase ^ 1);
                ptx::mbar_arrive_expect_tx(v_full(stage), Cfg::KV_BYTES);
                for (int c = 0; c < CH; ++c) tma_load_4d(sv + c * Cfg::CHUNK_BYTES, &tmV, v_full(stage), c * 32, j * TK, hk, b);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ---------------------------------------------------------------- MMA issuer
            constexpr uint32_t idesc_qk = idesc_bf16(TQ, TK, 0);
            constexpr uint32_t idesc_pv = idesc_bf16(TQ, HD, 1);
            auto issue_qk = [&](int j, int stage, uint32_t phase) {
                const uint32_t sk = sKV + stage * 2 * Cfg::KV_BYTES;
                ptx::mbar_wait(k_full(stage), phase);
                ptx::tc_fence_after();
                const uint32_t d = tS0 + (j & 1) * 128;
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) {
                    const uint32_t off = (k >> 1) * Cfg::CHUNK_BYTES + (k & 1) * 32;
                    ptx::umma_bf16(d, desc_kmajor_sw64(sQ + off), desc_kmajor_sw64(sk + off), idesc_qk, k > 0 ? 1u : 0u);
                }
                ptx::umma_commit(k_empty(stage));
                ptx::umma_commit(s_full(j & 1));
            };
            ptx::mbar_wait(q_full, 0);
            ptx::tc_fence_after();
            int qk_stage = 0, pv_stage = 0;
            uint32_t qk_phase = 0, pv_phase = 0;
            issue_qk(0, qk_stage, qk_phase);
            if (++qk_stage == STAGES) { qk_stage = 0; qk_phase ^= 1; }
            for (int j = 0; j < n_tiles; ++j) {
                if (j + 1 < n_tiles) {
                    issue_qk(j + 1, qk_stage, qk_phase);
                    if (++qk_stage == STAGES) { qk_stage = 0; qk_phase ^= 1; }
                }
                // O += P_j V_j
                const uint32_t sv = sKV + pv_stage * 2 * Cfg::KV_BYTES + Cfg::KV_BYTES;
                ptx::mbar_wait(v_full(pv_stage), pv_phase);
                ptx::mbar_wait(p_ready(j & 1), (j >> 1) & 1);
                ptx::tc_fence_after();
                const uint32_t tP = tS0 + (j & 1) * 128;
#pragma unroll
                for (int k = 0; k < TK / 16; ++k) {
                    ptx::umma_bf16_ts(tO, tP + k * 8, desc_mnmajor_sw64(sv + k * 1024), idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
                }
                ptx::umma_commit(v_empty(pv_stage));
                ptx::umma_commit(o_done);
                if (++pv_stage == STAGES) { pv_stage = 0; pv_phase ^= 1; }
            }
        }
    } else {
        // ---------------------------------------------------------------- softmax / correction / epilogue
        const int q = warp & 3;
        const int row = m0 + q * 32 + lane;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        float m_used = -INFINITY, l_sum = 0.f;
        for (int j = 0; j < n_tiles; ++j) {
            const uint32_t tS = tS0 + (j & 1) * 128 + lane_off;
            ptx::mbar_wait(s_full(j & 1), (j >> 1) & 1);
            ptx::tc_fence_after();
            const int n0 = j * TK;
            const bool need_mask = (n0 + TK > p.skv) || (CAUSAL && (n0 + TK - 1 > m0 + q * 32 + causal_off));
            const int col_lim = CAUSAL ? min(p.skv, row + causal_off + 1) : p.skv;   // valid cols: [0, col_lim)
            // pass 1: row max of this tile. The masked variant is a separate instantiation so that full tiles
            // (all but the last / diagonal ones) carry no compare/select instructions at all.
            auto pass1 = [&](auto mask_tag) -> float {
                constexpr bool MASK = decltype(mask_tag)::value;
                float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    uint32_t r[32];
                    ptx::tmem_ld_32x32(tS + c * 32, r);
                    ptx::tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        float x0 = __uint_as_float(r[i]), x1 = __uint_as_float(r[i + 1]);
                        if (ROUND) { x0 = bf16r(x0); x1 = bf16r(x1); }
                        if (MASK) {
                            if (n0 + c * 32 + i >= col_lim) x0 = -INFINITY;
                            if (n0 + c * 32 + i + 1 >= col_lim) x1 = -INFINITY;
                        }
                        mx0 = fmaxf(mx0, x0);
                        mx1 = fmaxf(mx1, x1);
                    }
                }
                return fmaxf(mx0, mx1);
            };
            const float mx = need_mask ? pass1(std::true_type{}) : pass1(std::false_type{});
            const float m_tile = mx * p.scale_log2;
            // lazy rescale: keep the stale reference max unless it would let P grow beyond 2^RESCALE_LOG2
            float factor = 1.0f;
            bool need = false;
            if (m_tile > m_used + RESCALE_LOG2 || m_used == -INFINITY) {
                if (m_tile != -INFINITY) {
                    need = (m_used != -INFINITY);
                    factor = need ? exp2f(m_used - m_tile) : 1.0f;
                    m_used = m_tile;
                }
            }
            if (__any_sync(0xffffffffu, need)) {
                // O must hold PV_{j-1} before it is scaled (at most one phase behind, see header)
                ptx::mbar_wait(o_done, (j - 1) & 1);
                ptx::tc_fence_after();
#pragma unroll 1
                for (int c = 0; c < HD / 32; ++c) {
                    uint32_t r[32];
                    ptx::tmem_ld_32x32(tO + lane_off + c * 32, r);
                    ptx::tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * factor);
                    // 2 x 16-column stores
                    uint32_t lo[16], hi[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) { lo[i] = r[i]; hi[i] = r[16 + i]; }
                    ptx::tmem_st_32x16(tO + lane_off + c * 32, lo);
                    ptx::tmem_st_32x16(tO + lane_off + c * 32 + 16, hi);
                }
                l_sum *= factor;
            }
            // pass 2: P = exp2(s * scale_log2 - m_used), row sum, bf16 P back into TMEM over S
            const float neg_m = (m_used == -INFINITY) ? 0.f : -m_used;
            auto pass2 = [&](auto mask_tag) -> float {
                constexpr bool MASK = decltype(mask_tag)::value;
                float acc0 = 0.f, acc1 = 0.f;
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    uint32_t r[32];
                    ptx::tmem_ld_32x32(tS + c * 32, r);
                    ptx::tmem_wait_ld();
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float x0 = __uint_as_float(r[2 * i]), x1 = __uint_as_float(r[2 * i + 1]);
                        if (ROUND) { x0 = bf16r(x0); x1 = bf16r(x1); }
                        float p0 = fast_exp2(fmaf(x0, p.scale_log2, neg_m));
                        float p1 = fast_exp2(fmaf(x1, p.scale_log2, neg_m));
                        if (MASK) {
                            if (n0 + c * 32 + 2 * i >= col_lim) p0 = 0.f;
                            if (n0 + c * 32 + 2 * i + 1 >= col_lim) p1 = 0.f;
                        }
                        acc0 += p0;
                        acc1 += p1;
                        pk[i] = pack_bf16(p0, p1);
                    }
                    ptx::tmem_st_32x16(tS + c * 16, pk);
                }
                return acc0 + acc1;
            };
            l_sum += need_mask ? pass2(std::true_type{}) : pass2(std::false_type{});
            ptx::tmem_wait_st();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(p_ready(j & 1));
        }
        // epilogue: O / l -> bf16 -> global
        ptx::mbar_wait(o_done, (n_tiles - 1) & 1);
        ptx::tc_fence_after();
        const float inv = l_sum > 0.f ? 1.0f / l_sum : 0.f;
        __nv_bfloat16* op = p.o + (long long)b * p.o_bs + (long long)h * p.o_hs + (long long)row * p.o_ts;
#pragma unroll 1
        for (int c = 0; c < HD / 32; ++c) {
            uint32_t r[32];
            ptx::tmem_ld_32x32(tO + lane_off + c * 32, r);
            ptx::tmem_wait_ld();
            if (row < p.sq) {
#pragma unroll
                for (int g8 = 0; g8 < 4; ++g8) {
                    if (c * 32 + g8 * 8 < p.o_dim) {
                        uint4 o;
                        o.x = pack_bf16(__uint_as_float(r[g8 * 8 + 0]) * inv, __uint_as_float(r[g8 * 8 + 1]) * inv);
                        o.y = pack_bf16(__uint_as_float(r[g8 * 8 + 2]) * inv, __uint_as_float(r[g8 * 8 + 3]) * inv);
                        o.z = pack_bf16(__uint_as_float(r[g8 * 8 + 4]) * inv, __uint_as_float(r[g8 * 8 + 5]) * inv);
                        o.w = pack_bf16(__uint_as_float(r[g8 * 8 + 6]) * inv, __uint_as_float(r[g8 * 8 + 7]) * inv);
                        *reinterpret_cast<uint4*>(op + c * 32 + g8 * 8) = o;
                    }
                }
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem, 512);
    }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

// [batch, tokens, heads, D] bf16 with element strides (bs, ts, hs); box = 32 elements x 128 tokens.
int make_tmap_4d(CUtensorMap* tm, const void* ptr, int D, int tokens, int heads, int batch, long long ts, long long hs,
                 long long bs) {
    PFN_encodeTiled fn = encode_fn();
    if (!fn) return GVL_ERR_DRIVER;
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ts * 2) % 16 || (hs * 2) % 16 || (bs * 2) % 16) return GVL_ERR_ALIGN;
    cuuint64_t gdim[4] = {(cuuint64_t)D, (cuuint64_t)tokens, (cuuint64_t)heads, (cuuint64_t)batch};
    long long hs_b = hs * 2, bs_b = bs * 2;
    if (heads == 1 && hs_b == 0) hs_b = 16;
    if (batch == 1 && bs_b == 0) bs_b = 16;
    cuuint64_t gstr[3] = {(cuuint64_t)(ts * 2), (cuuint64_t)hs_b, (cuuint64_t)bs_b};
    cuuint32_t box[4] = {32, 128, 1, 1};
    cuuint32_t est[4] = {1, 1, 1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim, gstr, box, est,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? GVL_OK : GVL_ERR_DRIVER;
}

template <int HD, bool CAUSAL, bool ROUND = false>
int launch_tc(const AttnArgs& a, cudaStream_t stream) {
    if (!ROUND && a.round_scores) return launch_tc<HD, CAUSAL, true>(a, stream);
    using Cfg = AtcCfg<HD>;
    CUtensorMap tq, tk, tv;
    int rc;
    if ((rc = make_tmap_4d(&tq, a.q, HD, a.sq, a.heads, a.batch, a.q_ts, a.q_hs, a.q_bs)) != GVL_OK) return rc;
    if ((rc = make_tmap_4d(&tk, a.k, HD, a.skv, a.kv_heads, a.batch, a.k_ts, a.k_hs, a.k_bs)) != GVL_OK) return rc;
    if ((rc = make_tmap_4d(&tv, a.v, HD, a.skv, a.kv_heads, a.batch, a.v_ts, a.v_hs, a.v_bs)) != GVL_OK) return rc;
    auto kern = attn_tc_kernel<HD, CAUSAL, ROUND>;
    static unsigned long long attr_devs = 0ull;
    if (first_use_on_device(attr_devs)) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM) != cudaSuccess) return GVL_ERR_CUDA;
    }
    AtcParams p;
    p.o = a.o; p.o_bs = a.o_bs; p.o_ts = a.o_ts; p.o_hs = a.o_hs;
    p.sq = a.sq; p.skv = a.skv; p.heads = a.heads; p.kv_heads = a.kv_heads;
    p.o_dim = a.o_dim > 0 ? a.o_dim : a.head_dim;
    p.scale_log2 = a.scale * 1.4426950408889634f;
    p.round_scores = a.round_scores;
    dim3 grid((a.sq + TQ - 1) / TQ, a.heads, a.batch);
    kern<<<grid, ATC_THREADS, Cfg::SMEM, stream>>>(tq, tk, tv, p);
    g_launch_count++;
    return cudaGetLastError() == cudaSuccess ? GVL_OK : GVL_ERR_CUDA;
}

}  // namespace

int make_tmap_4d_attn(CUtensorMap* tm, const void* ptr, int D, int tokens, int heads, int batch, long long ts, long long hs,
                      long long bs) {
    return make_tmap_4d(tm, ptr, D, tokens, heads, batch, ts, hs, bs);
}

bool attention_tc_supported(const AttnArgs& a) {
    if (a.head_dim != 64 && a.head_dim != 96 && a.head_dim != 128) return false;
    if (a.o_dim % 8 != 0) return false;
    auto al = [](long long s) { return (s * 2) % 16 == 0; };
    return al(a.q_ts) && al(a.q_hs) && al(a.q_bs) && al(a.k_ts) && al(a.k_hs) && al(a.k_bs) && al(a.v_ts) && al(a.v_hs) &&
           al(a.v_bs) && al(a.o_ts) && al(a.o_hs) && al(a.o_bs) && a.sq >= 1 && a.skv >= 1 && a.skv >= a.sq * (a.causal ? 1 : 0);
}

int attention_tc_fwd(const AttnArgs& a, cudaStream_t stream) {
    if (a.causal) {
        if (a.head_dim == 64) return launch_tc<64, true>(a, stream);
        if (a.head_dim == 96) return launch_tc<96, true>(a, stream);
        return launch_tc<128, true>(a, stream);
    }
    if (a.head_dim == 64) return launch_tc<64, false>(a, stream);
    if (a.head_dim == 96) return launch_tc<96, false>(a, stream);
    return launch_tc<128, false>(a, stream);
}

}  // namespace gvl
