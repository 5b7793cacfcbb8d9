// tcgen05 attention, second generation: TWO query tiles (2 x 128 rows) per CTA, each with its own softmax
// warpgroup, sharing every K/V tile that TMA brings in. Motivation (profiles/r1_attention_tc.md): with one softmax
// warp per SM sub-partition the first kernel exposed every TMEM-load / MUFU latency and sat at ~3000 clk per
// 128x128 tile against a ~1024 clk MUFU.EX2 bound; two warps per sub-partition (one per query tile) overlap each
// other's stalls, K/V smem traffic per query row halves, and the tensor pipe always has the other tile's QK / PV.
//
//   warp 0        TMA producer (Q_A, Q_B once; K_j, V_j ring; SWIZZLE_64B chunks of [128 rows x 64 B])
//   warp 1        MMA issuer: QK_A(0) QK_B(0) | PV_A(j) QK_A(j+1) PV_B(j) QK_B(j+1) ...
//   warps 2..5    softmax / correction / epilogue of query tile A   (TMEM lanes = rows)
//   warps 6..9    same for query tile B
// TMEM: S_A [0,128) S_B [128,256) O_A [256,256+hd) O_B [384,384+hd); P_g (bf16) aliases S_g[0,64).
// Numerics identical to attention_tc.cu.
#include <cuda.h>
#include <type_traits>
#include "gvl_internal.h"
#include "ptx.cuh"

namespace gvl {

int make_tmap_4d_attn(CUtensorMap* tm, const void* ptr, int D, int tokens, int heads, int batch, long long ts, long long hs,
                      long long bs);

namespace {

constexpr int T2Q = 128;
constexpr int T2K = 128;
constexpr int T2_THREADS = 320;
constexpr float T2_RESCALE_LOG2 = 8.0f;

template <int HD>
struct Atc2Cfg {
    static constexpr int CH = HD / 32;
    static constexpr int Q_BYTES = T2Q * HD * 2;              // one query tile
    static constexpr int KV_BYTES = T2K * HD * 2;
    static constexpr int STAGES = (HD <= 96) ? 3 : 2;
    static constexpr int SMEM = 2 * Q_BYTES + STAGES * 2 * KV_BYTES + 1024 + 256;
    static constexpr int CHUNK_BYTES = 128 * 64;
};

__device__ __forceinline__ uint64_t d_kmajor_sw64(uint32_t addr) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)4 << 61);
}
__device__ __forceinline__ uint64_t d_mnmajor_sw64(uint32_t addr) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)(8192 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}
__host__ __device__ constexpr uint32_t idesc2(int M, int N, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(b_mn_major) << 16) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}
__device__ __forceinline__ void tma4(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

struct Atc2Params {
    __nv_bfloat16* o;
    long long o_bs, o_ts, o_hs;
    int sq, skv, heads, kv_heads, o_dim;
    float scale_log2;
};

template <int HD, bool CAUSAL, bool ROUND>
__global__ void __launch_bounds__(T2_THREADS, 1)
attn_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const Atc2Params p) {
    using Cfg = Atc2Cfg<HD>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int CH = Cfg::CH;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sQ = base;                                   // tile g at sQ + g*Q_BYTES
    const uint32_t sKV = base + 2 * Cfg::Q_BYTES;
    const uint32_t bars = sKV + STAGES * 2 * Cfg::KV_BYTES;
    const uint32_t q_full = bars;
    auto k_full = [&](int s) { return bars + 8u * (1 + s); };
    auto v_full = [&](int s) { return bars + 8u * (1 + STAGES + s); };
    auto k_empty = [&](int s) { return bars + 8u * (1 + 2 * STAGES + s); };
    auto v_empty = [&](int s) { return bars + 8u * (1 + 3 * STAGES + s); };
    auto s_full = [&](int g) { return bars + 8u * (1 + 4 * STAGES + g); };
    auto p_ready = [&](int g) { return bars + 8u * (3 + 4 * STAGES + g); };
    auto o_done = [&](int g) { return bars + 8u * (5 + 4 * STAGES + g); };
    const uint32_t tmem_slot = bars + 8u * (7 + 4 * STAGES);
    uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * 2 * T2Q;
    const int h = blockIdx.y, b = blockIdx.z;
    const int hk = h / (p.heads / p.kv_heads);
    const int causal_off = p.skv - p.sq;
    // KV tiles each query tile needs
    int n_g[2];
#pragma unroll
    for (int g = 0; g < 2; ++g) {
        const int r0 = m0 + g * T2Q;
        if (r0 >= p.sq) { n_g[g] = 0; continue; }
        int kv_end = p.skv;
        if (CAUSAL) {
            const int last = r0 + T2Q + causal_off;
            kv_end = last < p.skv ? last : p.skv;
            if (kv_end < 1) kv_end = 1;
        }
        n_g[g] = (kv_end + T2K - 1) / T2K;
    }
    const int n_tiles = n_g[0] > n_g[1] ? n_g[0] : n_g[1];

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
        for (int g = 0; g < 2; ++g) {
            ptx::mbar_init(s_full(g), 1);
            ptx::mbar_init(p_ready(g), 4);
            ptx::mbar_init(o_done(g), 1);
        }
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

    if (warp == 0) {
        if (lane == 0) {
            // ---------------------------------------------------------------- TMA producer
            ptx::mbar_arrive_expect_tx(q_full, 2 * Cfg::Q_BYTES);
            for (int g = 0; g < 2; ++g)
                for (int c = 0; c < CH; ++c)
                    tma4(sQ + g * Cfg::Q_BYTES + c * Cfg::CHUNK_BYTES, &tmQ, q_full, c * 32, m0 + g * T2Q, h, b);
            int stage = 0;
            uint32_t phase = 0;
            for (int j = 0; j < n_tiles; ++j) {
                const uint32_t sk = sKV + stage * 2 * Cfg::KV_BYTES;
                const uint32_t sv = sk + Cfg::KV_BYTES;
                ptx::mbar_wait(k_empty(stage), phase ^ 1);
                ptx::mbar_arrive_expect_tx(k_full(stage), Cfg::KV_BYTES);
                for (int c = 0; c < CH; ++c) tma4(sk + c * Cfg::CHUNK_BYTES, &tmK, k_full(stage), c * 32, j * T2K, hk, b);
                ptx::mbar_wait(v_empty(stage), phase ^ 1);
                ptx::mbar_arrive_expect_tx(v_full(stage), Cfg::KV_BYTES);
                for (int c = 0; c < CH; ++c) tma4(sv + c * Cfg::CHUNK_BYTES, &tmV, v_full(stage), c * 32, j * T2K, hk, b);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ---------------------------------------------------------------- MMA issuer
            constexpr uint32_t idesc_qk = idesc2(T2Q, T2K, 0);
            constexpr uint32_t idesc_pv = idesc2(T2Q, HD, 1);
            auto qk = [&](int g, uint32_t sk) {
                const uint32_t d = tmem + g * 128;
                const uint32_t sq_ = sQ + g * Cfg::Q_BYTES;
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) {
                    const uint32_t off = (k >> 1) * Cfg::CHUNK_BYTES + (k & 1) * 32;
                    ptx::umma_bf16(d, d_kmajor_sw64(sq_ + off), d_kmajor_sw64(sk + off), idesc_qk, k > 0 ? 1u : 0u);
                }
                ptx::umma_commit(s_full(g));
            };
            auto pv = [&](int g, int j, uint32_t sv) {
                ptx::mbar_wait(p_ready(g), j & 1);
                ptx::tc_fence_after();
                const uint32_t tP = tmem + g * 128;
                const uint32_t tO = tmem + 256 + g * 128;
#pragma unroll
                for (int k = 0; k < T2K / 16; ++k)
                    ptx::umma_bf16_ts(tO, tP + k * 8, d_mnmajor_sw64(sv + k * 1024), idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
                ptx::umma_commit(o_done(g));
            };
            ptx::mbar_wait(q_full, 0);
            int ks = 0, vs = 0;            // ring stages of the K tile of QK(j+1) and the V tile of PV(j)
            uint32_t kph = 0, vph = 0;
            if (n_tiles > 0) {
                ptx::mbar_wait(k_full(0), 0);
                ptx::tc_fence_after();
                if (n_g[0] > 0) qk(0, sKV);
                if (n_g[1] > 0) qk(1, sKV);
                ptx::umma_commit(k_empty(0));
                ks = 1 % STAGES;
                kph = (STAGES == 1) ? 1 : 0;
            }
            for (int j = 0; j < n_tiles; ++j) {
                const bool next = j + 1 < n_tiles;
                const uint32_t sv = sKV + vs * 2 * Cfg::KV_BYTES + Cfg::KV_BYTES;
                const uint32_t skn = sKV + ks * 2 * Cfg::KV_BYTES;
                ptx::mbar_wait(v_full(vs), vph);
                if (next) ptx::mbar_wait(k_full(ks), kph);
                ptx::tc_fence_after();
                if (j < n_g[0]) pv(0, j, sv);
                if (next && j + 1 < n_g[0]) qk(0, skn);
                if (j < n_g[1]) pv(1, j, sv);
                ptx::umma_commit(v_empty(vs));
                if (next && j + 1 < n_g[1]) qk(1, skn);
                if (next) {
                    ptx::umma_commit(k_empty(ks));
                    if (++ks == STAGES) { ks = 0; kph ^= 1; }
                }
                if (++vs == STAGES) { vs = 0; vph ^= 1; }
            }
        }
    } else {
        // ---------------------------------------------------------------- softmax / correction / epilogue
        const int g = (warp - 2) >> 2;                 // query tile of this warpgroup
        const int q = warp & 3;                        // TMEM lane quarter
        const int my_tiles = n_g[g];
        if (my_tiles > 0) {
            const int row0 = m0 + g * T2Q;
            const int row = row0 + q * 32 + lane;
            const uint32_t lane_off = (uint32_t)(q * 32) << 16;
            const uint32_t tS = tmem + g * 128 + lane_off;
            const uint32_t tO = tmem + 256 + g * 128 + lane_off;
            float m_used = -INFINITY, l_sum = 0.f;
            for (int j = 0; j < my_tiles; ++j) {
                ptx::mbar_wait(s_full(g), j & 1);
                ptx::tc_fence_after();
                const int n0 = j * T2K;
                const bool need_mask = (n0 + T2K > p.skv) || (CAUSAL && (n0 + T2K - 1 > row0 + q * 32 + causal_off));
                const int col_lim = CAUSAL ? min(p.skv, row + causal_off + 1) : p.skv;
                auto pass1 = [&](auto mask_tag) -> float {
                    constexpr bool MASK = decltype(mask_tag)::value;
                    float mx0 = -INFINITY, mx1 = -INFINITY;
                    // software-pipelined TMEM reads: the load of chunk c+1 is in flight while chunk c is reduced
                    uint32_t rr[2][32];
                    ptx::tmem_ld_32x32(tS, rr[0]);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        ptx::tmem_wait_ld();
                        if (c + 1 < 4) ptx::tmem_ld_32x32(tS + (c + 1) * 32, rr[(c + 1) & 1]);
                        const uint32_t (&r)[32] = rr[c & 1];
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
                float factor = 1.0f;
                bool need = false;
                if (m_tile > m_used + T2_RESCALE_LOG2 || m_used == -INFINITY) {
                    if (m_tile != -INFINITY) {
                        need = (m_used != -INFINITY);
                        factor = need ? exp2f(m_used - m_tile) : 1.0f;
                        m_used = m_tile;
                    }
                }
                if (__any_sync(0xffffffffu, need)) {
                    ptx::mbar_wait(o_done(g), (j - 1) & 1);
                    ptx::tc_fence_after();
#pragma unroll 1
                    for (int c = 0; c < HD / 32; ++c) {
                        uint32_t r[32];
                        ptx::tmem_ld_32x32(tO + c * 32, r);
                        ptx::tmem_wait_ld();
                        uint32_t lo[16], hi[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            lo[i] = __float_as_uint(__uint_as_float(r[i]) * factor);
                            hi[i] = __float_as_uint(__uint_as_float(r[16 + i]) * factor);
                        }
                        ptx::tmem_st_32x16(tO + c * 32, lo);
                        ptx::tmem_st_32x16(tO + c * 32 + 16, hi);
                    }
                    l_sum *= factor;
                }
                const float neg_m = (m_used == -INFINITY) ? 0.f : -m_used;
                auto pass2 = [&](auto mask_tag) -> float {
                    constexpr bool MASK = decltype(mask_tag)::value;
                    float acc0 = 0.f, acc1 = 0.f;
                    uint32_t rr[2][32];
                    ptx::tmem_ld_32x32(tS, rr[0]);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        ptx::tmem_wait_ld();
                        if (c + 1 < 4) ptx::tmem_ld_32x32(tS + (c + 1) * 32, rr[(c + 1) & 1]);
                        const uint32_t (&r)[32] = rr[c & 1];
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
                        // P chunk c lands on columns [16c, 16c+16) = S columns already consumed (chunks <= c/2)
                        ptx::tmem_st_32x16(tS + c * 16, pk);
                    }
                    return acc0 + acc1;
                };
                l_sum += need_mask ? pass2(std::true_type{}) : pass2(std::false_type{});
                ptx::tmem_wait_st();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(p_ready(g));
            }
            // epilogue: O / l -> bf16 -> global
            ptx::mbar_wait(o_done(g), (my_tiles - 1) & 1);
            ptx::tc_fence_after();
            const float inv = l_sum > 0.f ? 1.0f / l_sum : 0.f;
            __nv_bfloat16* op = p.o + (long long)b * p.o_bs + (long long)h * p.o_hs + (long long)row * p.o_ts;
#pragma unroll 1
            for (int c = 0; c < HD / 32; ++c) {
                uint32_t r[32];
                ptx::tmem_ld_32x32(tO + c * 32, r);
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
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem, 512);
    }
}

template <int HD, bool CAUSAL, bool ROUND = false>
int launch_tc2(const AttnArgs& a, cudaStream_t stream) {
    if (!ROUND && a.round_scores) return launch_tc2<HD, CAUSAL, true>(a, stream);
    using Cfg = Atc2Cfg<HD>;
    CUtensorMap tq, tk, tv;
    int rc;
    if ((rc = make_tmap_4d_attn(&tq, a.q, HD, a.sq, a.heads, a.batch, a.q_ts, a.q_hs, a.q_bs)) != GVL_OK) return rc;
    if ((rc = make_tmap_4d_attn(&tk, a.k, HD, a.skv, a.kv_heads, a.batch, a.k_ts, a.k_hs, a.k_bs)) != GVL_OK) return rc;
    if ((rc = make_tmap_4d_attn(&tv, a.v, HD, a.skv, a.kv_heads, a.batch, a.v_ts, a.v_hs, a.v_bs)) != GVL_OK) return rc;
    auto kern = attn_tc2_kernel<HD, CAUSAL, ROUND>;
    static unsigned long long attr_devs = 0ull;
    if (first_use_on_device(attr_devs)) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM) != cudaSuccess) return GVL_ERR_CUDA;
    }
    Atc2Params p;
    p.o = a.o; p.o_bs = a.o_bs; p.o_ts = a.o_ts; p.o_hs = a.o_hs;
    p.sq = a.sq; p.skv = a.skv; p.heads = a.heads; p.kv_heads = a.kv_heads;
    p.o_dim = a.o_dim > 0 ? a.o_dim : a.head_dim;
    p.scale_log2 = a.scale * 1.4426950408889634f;
    dim3 grid((a.sq + 2 * T2Q - 1) / (2 * T2Q), a.heads, a.batch);
    kern<<<grid, T2_THREADS, Cfg::SMEM, stream>>>(tq, tk, tv, p);
    g_launch_count++;
    return cudaGetLastError() == cudaSuccess ? GVL_OK : GVL_ERR_CUDA;
}

}  // namespace

int attention_tc2_fwd(const AttnArgs& a, cudaStream_t stream) {
    if (a.causal) {
        if (a.head_dim == 64) return launch_tc2<64, true>(a, stream);
        if (a.head_dim == 96) return launch_tc2<96, true>(a, stream);
        return launch_tc2<128, true>(a, stream);
    }
    if (a.head_dim == 64) return launch_tc2<64, false>(a, stream);
    if (a.head_dim == 96) return launch_tc2<96, false>(a, stream);
    return launch_tc2<128, false>(a, stream);
}

}  // namespace gvl
