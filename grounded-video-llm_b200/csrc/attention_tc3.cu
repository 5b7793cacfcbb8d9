// tcgen05 attention, third generation: two query tiles per CTA (as attention_tc2.cu) but the score tile of each softmax
// step is 128 rows x 64 keys and DOUBLE-BUFFERED in TMEM, so QK of step t+1 is already finished while the softmax
// warpgroup works on step t. In the second generation S and P aliased a single 128-column buffer per query tile: the
// next QK could only be issued after PV had consumed P, so every softmax warpgroup idled for QK + PV (~770 clk) per
// 128x128 tile against ~1024 clk of MUFU work (profiles/r1_attention_tc.md) -- about 2x off the exp2 bound.
//
//   warp 0        TMA producer (Q_A, Q_B once; K_j, V_j ring of 128-key tiles; SWIZZLE_64B chunks of [128 rows x 64 B])
//   warp 1 / 10   MMA issuer of query tile A / B: QK(g,0) QK(g,1) | PV(g,t) QK(g,t+2) ...   t = 64-key step. One issuer
//                 per tile: a single thread needs ~100 clk of descriptor arithmetic per tcgen05.mma, and 20 MMAs per step
//                 for both tiles made the issuer, not MUFU.EX2, the bottleneck (profiles/r1_attention_tc.md)
//   warps 2..5    softmax / correction / epilogue of query tile A   (TMEM lanes = rows, ONE pass: 64 scores in registers)
//   warps 6..9    same for query tile B
// TMEM (512 columns): O_A [0,128) O_B [128,256) S_{g,b} [256 + (2g+b)*64, +64), g = tile, b = step parity;
// P_{g,b} (bf16) overwrites the first 32 columns of S_{g,b} once the owning thread holds the 64 scores in registers.
// Numerics identical to attention_tc.cu / attention_tc2.cu (lazy rescale, bf16 P, fp32 O).
#include <cuda.h>
#include <type_traits>
#include <stdlib.h>
#include "gvl_internal.h"
#include "ptx.cuh"

namespace gvl {

int make_tmap_4d_attn(CUtensorMap* tm, const void* ptr, int D, int tokens, int heads, int batch, long long ts, long long hs,
                      long long bs);

namespace {

constexpr int T3Q = 128;          // query rows per tile
constexpr int T3K = 128;          // keys per shared-memory stage (one TMA tile)
constexpr int T3S = 64;           // keys per softmax step (columns of one S buffer)
constexpr int T3_THREADS = 352;   // TMA, MMA issuer A, 4 + 4 softmax warps, MMA issuer B
constexpr float T3_RESCALE_LOG2 = 8.0f;

template <int HD>
struct Atc3Cfg {
    static constexpr int CH = HD / 32;
    static constexpr int Q_BYTES = T3Q * HD * 2;
    static constexpr int KV_BYTES = T3K * HD * 2;
    static constexpr int STAGES = (HD <= 96) ? 3 : 2;
    static constexpr int SMEM = 2 * Q_BYTES + STAGES * 2 * KV_BYTES + 1024 + 256;
    static constexpr int CHUNK_BYTES = 128 * 64;
};

__device__ __forceinline__ uint64_t d3_kmajor_sw64(uint32_t addr) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)4 << 61);
}
__device__ __forceinline__ uint64_t d3_mnmajor_sw64(uint32_t addr) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)(8192 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}
__host__ __device__ constexpr uint32_t idesc3(int M, int N, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(b_mn_major) << 16) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}
__device__ __forceinline__ void tma4_3(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// wait with back-off for warps that are NOT on the latency-critical chain (the TMA producer runs stages ahead): a bare
// try_wait spin competes for issue slots with the softmax warps of the same scheduler
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity) {
    while (!ptx::mbar_try_wait(bar, parity)) __nanosleep(64);
}
__device__ __forceinline__ float max3f(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
struct Atc3Params {
    __nv_bfloat16* o;
    long long o_bs, o_ts, o_hs;
    int sq, skv, heads, kv_heads, o_dim;
    float scale_log2;
    int variant;
};

template <int HD, bool CAUSAL, bool ROUND>
__global__ void __launch_bounds__(T3_THREADS, 1)
attn_tc3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const Atc3Params p) {
    using Cfg = Atc3Cfg<HD>;
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
    auto s_full = [&](int g, int b) { return bars + 8u * (1 + 4 * STAGES + g * 2 + b); };
    auto p_ready = [&](int g, int b) { return bars + 8u * (5 + 4 * STAGES + g * 2 + b); };
    auto o_done = [&](int g) { return bars + 8u * (9 + 4 * STAGES + g); };
    // completion of the LAST PV of a tile. o_done flips once per step and may lag the softmax by two steps now that S is
    // double-buffered, so its parity cannot tell "all done" from "two steps behind" at the epilogue.
    auto o_final = [&](int g) { return bars + 8u * (11 + 4 * STAGES + g); };
    const uint32_t tmem_slot = bars + 8u * (13 + 4 * STAGES);
    uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * 2 * T3Q;
    const int h = blockIdx.y, b = blockIdx.z;
    const int hk = h / (p.heads / p.kv_heads);
    const int causal_off = p.skv - p.sq;
    // 64-key softmax steps each query tile needs
    int n_sub[2];
#pragma unroll
    for (int g = 0; g < 2; ++g) {
        const int r0 = m0 + g * T3Q;
        if (r0 >= p.sq) { n_sub[g] = 0; continue; }
        int kv_end = p.skv;
        if (CAUSAL) {
            const int last = r0 + T3Q + causal_off;
            kv_end = last < p.skv ? last : p.skv;
            if (kv_end < 1) kv_end = 1;
        }
        n_sub[g] = (kv_end + T3S - 1) / T3S;
    }
    const int max_sub = n_sub[0] > n_sub[1] ? n_sub[0] : n_sub[1];
    const int n_tiles = (max_sub + 1) / 2;                      // 128-key tiles to bring in

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmQ);
        ptx::prefetch_tmap(&tmK);
        ptx::prefetch_tmap(&tmV);
        ptx::mbar_init(q_full, 1);
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(k_full(s), 1);
            ptx::mbar_init(v_full(s), 1);
            ptx::mbar_init(k_empty(s), 2);                       // both issuers release a stage
            ptx::mbar_init(v_empty(s), 2);
        }
        for (int g = 0; g < 2; ++g) {
            for (int bb = 0; bb < 2; ++bb) {
                ptx::mbar_init(s_full(g, bb), 1);
                ptx::mbar_init(p_ready(g, bb), 4);
            }
            ptx::mbar_init(o_done(g), 1);
            ptx::mbar_init(o_final(g), 1);
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
                    tma4_3(sQ + g * Cfg::Q_BYTES + c * Cfg::CHUNK_BYTES, &tmQ, q_full, c * 32, m0 + g * T3Q, h, b);
            int stage = 0;
            uint32_t phase = 0;
            for (int j = 0; j < n_tiles; ++j) {
                const uint32_t sk = sKV + stage * 2 * Cfg::KV_BYTES;
                const uint32_t sv = sk + Cfg::KV_BYTES;
                mbar_wait_backoff(k_empty(stage), phase ^ 1);
                ptx::mbar_arrive_expect_tx(k_full(stage), Cfg::KV_BYTES);
                for (int c = 0; c < CH; ++c) tma4_3(sk + c * Cfg::CHUNK_BYTES, &tmK, k_full(stage), c * 32, j * T3K, hk, b);
                mbar_wait_backoff(v_empty(stage), phase ^ 1);
                ptx::mbar_arrive_expect_tx(v_full(stage), Cfg::KV_BYTES);
                for (int c = 0; c < CH; ++c) tma4_3(sv + c * Cfg::CHUNK_BYTES, &tmV, v_full(stage), c * 32, j * T3K, hk, b);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1 || warp == 10) {
        {   // the whole warp walks the loop (uniform control flow, descriptors stay in uniform registers); one elected
            // lane issues the tcgen05 instructions
            // ---------------------------------------------------------------- MMA issuer of query tile g
            const int g = warp == 1 ? 0 : 1;
            const int my_sub = n_sub[g];
            constexpr uint32_t idesc_qk = idesc3(T3Q, T3S, 0);
            constexpr uint32_t idesc_pv = idesc3(T3Q, HD, 1);
            const uint32_t sq_ = sQ + g * Cfg::Q_BYTES;
            const uint32_t tO = tmem + g * 128;
            // S_{g, t&1} = Q_g K[64-key half t&1 of the stage]^T
            auto qk = [&](int t, uint32_t sk) {
                const uint32_t d = tmem + 256 + (g * 2 + (t & 1)) * T3S;
                const uint32_t skh = sk + (t & 1) * (T3S * 64);           // 64 rows x 64 B inside every chunk
                if (ptx::elect_one()) {
#pragma unroll
                    for (int k = 0; k < HD / 16; ++k) {
                        const uint32_t off = (k >> 1) * Cfg::CHUNK_BYTES + (k & 1) * 32;
                        ptx::umma_bf16(d, d3_kmajor_sw64(sq_ + off), d3_kmajor_sw64(skh + off), idesc_qk, k > 0 ? 1u : 0u);
                    }
                    ptx::umma_commit(s_full(g, t & 1));
                }
                __syncwarp();
            };
            // O_g += P_{g, t&1} V[64-key half]
            auto pv = [&](int t, uint32_t sv) {
                ptx::mbar_wait(p_ready(g, t & 1), (t >> 1) & 1);
                ptx::tc_fence_after();
                const uint32_t tP = tmem + 256 + (g * 2 + (t & 1)) * T3S;
                if (ptx::elect_one()) {
#pragma unroll
                    for (int k = 0; k < T3S / 16; ++k)
                        ptx::umma_bf16_ts(tO, tP + k * 8, d3_mnmajor_sw64(sv + ((t & 1) * (T3S / 16) + k) * 1024), idesc_pv,
                                          (t > 0 || k > 0) ? 1u : 0u);
                    ptx::umma_commit(o_done(g));
                    if (t == my_sub - 1) ptx::umma_commit(o_final(g));
                }
                __syncwarp();
            };
            ptx::mbar_wait(q_full, 0);
            int ks = 0, vs = 0;            // ring stages of the K tile of QK(2j+2..3) and the V tile of PV(2j..2j+1)
            uint32_t kph = 0, vph = 0;
            if (n_tiles > 0) {
                ptx::mbar_wait(k_full(0), 0);
                ptx::tc_fence_after();
                if (0 < my_sub) qk(0, sKV);
                if (1 < my_sub) qk(1, sKV);
                if (ptx::elect_one()) ptx::umma_commit(k_empty(0));
                __syncwarp();
                ks = 1 % STAGES;
                kph = (STAGES == 1) ? 1 : 0;
            }
            // every stage is released by BOTH issuers, also by the one whose tile needs fewer keys (causal) or is empty
            for (int j = 0; j < n_tiles; ++j) {
                const bool next = j + 1 < n_tiles;
                const uint32_t sv = sKV + vs * 2 * Cfg::KV_BYTES + Cfg::KV_BYTES;
                const uint32_t skn = sKV + ks * 2 * Cfg::KV_BYTES;
                ptx::mbar_wait(v_full(vs), vph);
                if (next) ptx::mbar_wait(k_full(ks), kph);
                ptx::tc_fence_after();
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int t = 2 * j + hh;
                    if (t < my_sub) pv(t, sv);
                    if (next && t + 2 < my_sub) qk(t + 2, skn);
                }
                if (ptx::elect_one()) {
                    ptx::umma_commit(v_empty(vs));
                    if (next) ptx::umma_commit(k_empty(ks));
                }
                __syncwarp();
                if (next) {
                    if (++ks == STAGES) { ks = 0; kph ^= 1; }
                }
                if (++vs == STAGES) { vs = 0; vph ^= 1; }
            }
        }
    } else {
        // ---------------------------------------------------------------- softmax / correction / epilogue
        const int g = (warp - 2) >> 2;                 // query tile of this warpgroup
        const int q = warp & 3;                        // TMEM lane quarter
        const int my_sub = n_sub[g];
        if (my_sub > 0) {
            const int row0 = m0 + g * T3Q;
            const int row = row0 + q * 32 + lane;
            const uint32_t lane_off = (uint32_t)(q * 32) << 16;
            const uint32_t tO = tmem + g * 128 + lane_off;
            float m_used = -INFINITY, l_sum = 0.f;
            const int col_lim = CAUSAL ? min(p.skv, row + causal_off + 1) : p.skv;
            long long tacc[6] = {0, 0, 0, 0, 0, 0};
            const bool timing = (p.variant & 32) && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (warp == 2 || warp == 6) && lane == 0;
            for (int t = 0; t < my_sub; ++t) {
                long long c0 = 0, c1 = 0;
                if (timing) c0 = clock64();
#define T3_MARK(i) do { if (timing) { c1 = clock64(); tacc[i] += c1 - c0; c0 = c1; } } while (0)
                const uint32_t tS = tmem + 256 + (g * 2 + (t & 1)) * T3S + lane_off;
                ptx::mbar_wait(s_full(g, t & 1), (t >> 1) & 1);
                ptx::tc_fence_after();
                T3_MARK(0);
                uint32_t r0[32], r1[32];
                ptx::tmem_ld_32x32(tS, r0);
                ptx::tmem_ld_32x32(tS + 32, r1);
                const int n0 = t * T3S;
                const bool need_mask = (n0 + T3S > p.skv) || (CAUSAL && (n0 + T3S - 1 > row0 + q * 32 + causal_off));
                ptx::tmem_wait_ld();
                T3_MARK(1);
                // ---- row maximum of the 64 scores (kept in registers: one TMEM pass per step)
                if (need_mask || ROUND) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        float x0 = __uint_as_float(r0[i]), x1 = __uint_as_float(r1[i]);
                        if (ROUND) { x0 = bf16r(x0); x1 = bf16r(x1); }
                        if (need_mask) {
                            if (n0 + i >= col_lim) x0 = -INFINITY;
                            if (n0 + 32 + i >= col_lim) x1 = -INFINITY;
                        }
                        r0[i] = __float_as_uint(x0);
                        r1[i] = __float_as_uint(x1);
                    }
                }
                // four independent 3-input max chains (8 deep) instead of two 32-deep ones
                float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    mx[0] = max3f(mx[0], __uint_as_float(r0[i]), __uint_as_float(r0[i + 1]));
                    mx[1] = max3f(mx[1], __uint_as_float(r0[16 + i]), __uint_as_float(r0[17 + i]));
                    mx[2] = max3f(mx[2], __uint_as_float(r1[i]), __uint_as_float(r1[i + 1]));
                    mx[3] = max3f(mx[3], __uint_as_float(r1[16 + i]), __uint_as_float(r1[17 + i]));
                }
                const float m_tile = max3f(mx[0], mx[1], fmaxf(mx[2], mx[3])) * p.scale_log2;
                bool need = false;
                float m_old = m_used;
                if (m_tile > m_used + T3_RESCALE_LOG2 || m_used == -INFINITY) {
                    if (m_tile != -INFINITY) {
                        need = (m_used != -INFINITY);
                        m_used = m_tile;
                    }
                }
                if (__any_sync(0xffffffffu, need)) {
                    const float factor = need ? fast_exp2(m_old - m_used) : 1.0f;
                    ptx::mbar_wait(o_done(g), (t - 1) & 1);
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
                T3_MARK(2);
                // ---- P = exp2(S * scale - m): masked scores are -inf and give exactly 0
                float acc0 = 0.f, acc1 = 0.f;
                uint32_t pk0[16], pk1[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float p0 = fast_exp2(fmaf(__uint_as_float(r0[2 * i]), p.scale_log2, neg_m));
                    const float p1 = fast_exp2(fmaf(__uint_as_float(r0[2 * i + 1]), p.scale_log2, neg_m));
                    const float p2 = fast_exp2(fmaf(__uint_as_float(r1[2 * i]), p.scale_log2, neg_m));
                    const float p3 = fast_exp2(fmaf(__uint_as_float(r1[2 * i + 1]), p.scale_log2, neg_m));
                    acc0 += p0 + p1;
                    acc1 += p2 + p3;
                    pk0[i] = pack_bf16(p0, p1);
                    pk1[i] = pack_bf16(p2, p3);
                }
                l_sum += acc0 + acc1;
                T3_MARK(3);
                ptx::tmem_st_32x16(tS, pk0);             // keys [0,32)  -> P columns [0,16)
                ptx::tmem_st_32x16(tS + 16, pk1);        // keys [32,64) -> P columns [16,32)
                ptx::tmem_wait_st();
                T3_MARK(4);
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(p_ready(g, t & 1));
                T3_MARK(5);
            }
#undef T3_MARK
            if (timing)
                printf("attn_tc3 timing warp %d steps %d: wait_S %lld  tmem_ld %lld  max+rescale %lld  exp+pack %lld  tmem_st %lld  arrive %lld (avg clk/step)\n",
                       warp, my_sub, tacc[0] / my_sub, tacc[1] / my_sub, tacc[2] / my_sub, tacc[3] / my_sub, tacc[4] / my_sub, tacc[5] / my_sub);
            // epilogue: O / l -> bf16 -> global
            ptx::mbar_wait(o_final(g), 0);
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
int launch_tc3(const AttnArgs& a, cudaStream_t stream) {
    if (!ROUND && a.round_scores) return launch_tc3<HD, CAUSAL, true>(a, stream);
    using Cfg = Atc3Cfg<HD>;
    CUtensorMap tq, tk, tv;
    int rc;
    if ((rc = make_tmap_4d_attn(&tq, a.q, HD, a.sq, a.heads, a.batch, a.q_ts, a.q_hs, a.q_bs)) != GVL_OK) return rc;
    if ((rc = make_tmap_4d_attn(&tk, a.k, HD, a.skv, a.kv_heads, a.batch, a.k_ts, a.k_hs, a.k_bs)) != GVL_OK) return rc;
    if ((rc = make_tmap_4d_attn(&tv, a.v, HD, a.skv, a.kv_heads, a.batch, a.v_ts, a.v_hs, a.v_bs)) != GVL_OK) return rc;
    auto kern = attn_tc3_kernel<HD, CAUSAL, ROUND>;
    static unsigned long long attr_devs = 0ull;
    if (first_use_on_device(attr_devs)) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM) != cudaSuccess) return GVL_ERR_CUDA;
    }
    Atc3Params p;
    p.o = a.o; p.o_bs = a.o_bs; p.o_ts = a.o_ts; p.o_hs = a.o_hs;
    p.sq = a.sq; p.skv = a.skv; p.heads = a.heads; p.kv_heads = a.kv_heads;
    p.o_dim = a.o_dim > 0 ? a.o_dim : a.head_dim;
    p.scale_log2 = a.scale * 1.4426950408889634f;
    static const int variant = getenv("GVL_ATTN_TIMING") ? 32 : 0;   // bring-up: per-step clock64 breakdown of one CTA's softmax warps
    p.variant = variant;
    dim3 grid((a.sq + 2 * T3Q - 1) / (2 * T3Q), a.heads, a.batch);
    kern<<<grid, T3_THREADS, Cfg::SMEM, stream>>>(tq, tk, tv, p);
    g_launch_count++;
    return cudaGetLastError() == cudaSuccess ? GVL_OK : GVL_ERR_CUDA;
}

}  // namespace

int attention_tc3_fwd(const AttnArgs& a, cudaStream_t stream) {
    if (a.causal) {
        if (a.head_dim == 64) return launch_tc3<64, true>(a, stream);
        if (a.head_dim == 96) return launch_tc3<96, true>(a, stream);
        return launch_tc3<128, true>(a, stream);
    }
    if (a.head_dim == 64) return launch_tc3<64, false>(a, stream);
    if (a.head_dim == 96) return launch_tc3<96, false>(a, stream);
    return launch_tc3<128, false>(a, stream);
}

}  // namespace gvl
