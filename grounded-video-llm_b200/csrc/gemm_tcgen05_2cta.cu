// 2-CTA (cta_group::2) variant of the tcgen05 GEMM: a CTA pair on one TPC computes a 256 x 256 output tile.
// Each CTA stages its own 128 rows of A and HALF of the W tile (128 of the 256 rows); one thread of the leader CTA
// issues tcgen05.mma.cta_group::2 (M=256, N=256, K=16), which reads A/W from both CTAs' shared memory and writes
// rows [0,128) of the accumulator into the leader's TMEM and rows [128,256) into the peer's. Per CTA and k-block the
// smem fill drops from 48 KB (1-CTA 128x256 tile) to 32 KB, i.e. 1.5x less L2->SM traffic per FLOP and room for
// 6 pipeline stages; that is the gap to the vendor GEMM the 1-CTA kernel left (profiles/r1_gemm.md).
//
// Synchronisation (per smem stage s / accumulator stage a):
//   full[s]    lives in the LEADER: both CTAs' TMA loads complete_tx on it (.cta_group::2 form, peer bit masked),
//              the leader's producer arrives once with expect_tx = 2 x 32 KB
//   empty[s]   one per CTA, armed by the leader's tcgen05.commit ... multicast::cluster (mask 0b11)
//   tfull[a]   one per CTA, multicast commit after the last k-block of a tile
//   tempty[a]  lives in the leader, 2 x 8 epilogue warps arrive (the peer's through mapa + mbarrier.arrive.cluster)
#include <cuda.h>
#include "gemm_epilogue.cuh"

namespace gvl {

int make_tmap_2d_bf16(CUtensorMap* tm, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                      uint32_t box_cols);

namespace {

constexpr int C2_BM = 128;          // rows of A per CTA (256 per pair)
constexpr int C2_BN = 256;          // tile N (each CTA stages 128 rows of W)
constexpr int C2_BK = 64;
constexpr int C2_STAGES = 6;
constexpr int C2_EPI_WARPS = 8;
constexpr int C2_THREADS = 64 + 32 * C2_EPI_WARPS;
constexpr int C2_A_BYTES = C2_BM * C2_BK * 2;           // 16 KB
constexpr int C2_B_BYTES = (C2_BN / 2) * C2_BK * 2;     // 16 KB
constexpr int C2_STAGE_BYTES = C2_A_BYTES + C2_B_BYTES;
constexpr int C2_SMEM = C2_STAGES * C2_STAGE_BYTES + 1024 + 256;
constexpr int C2_TMEM_COLS = 2 * C2_BN;

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
// arrive on the leader's tempty barrier. What it orders is this warp's tcgen05.ld of the accumulator stage (completed by
// tcgen05.wait::ld, fenced by tcgen05.fence::before_thread_sync) against the leader's next tcgen05.mma -- TMEM, not memory -- so
// the default (.release.cta) form is enough; the .release.cluster form compiled to MEMBAR.ALL.GPU + ERRBAR per tile and warp,
// ~10 % of the epilogue warps' stall samples in the IV2 fc1 capture (profiles/r2_gemm.md).
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t smem_dst, const void* tmap, uint32_t leader_bar, int c_inner,
                                                int c_outer) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(leader_bar), "r"(c_inner), "r"(c_outer)
        : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2cta() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

template <int ACT, int RES, bool OUT_F32>
__global__ void __launch_bounds__(C2_THREADS, 1)
gemm_bf16_tcgen05_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                              const GemmParams2 p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + C2_STAGES * C2_STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (C2_STAGES + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * C2_STAGES + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * C2_STAGES + 2 + s); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * C2_STAGES + 4);
    uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(smem_gen + C2_STAGES * C2_STAGE_BYTES + 8 * (2 * C2_STAGES + 4));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1;
    const int num_clusters = gridDim.x >> 1;
    const int num_kb = (p.K + C2_BK - 1) / C2_BK;
    const int num_tiles = p.num_m_tiles * p.num_n_tiles;   // tiles of 256 x 256

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmA);
        ptx::prefetch_tmap(&tmB);
        for (int s = 0; s < C2_STAGES; ++s) {
            ptx::mbar_init(full_bar(s), 1);
            ptx::mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(tfull_bar(s), 1);
            ptx::mbar_init(tempty_bar(s), 2 * C2_EPI_WARPS);
        }
        ptx::fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc_2cta(tmem_slot, C2_TMEM_COLS);
        tmem_relinquish_2cta();
    }
    ptx::tc_fence_before();
    cluster_sync_all();       // barriers of BOTH CTAs are initialised before anyone signals across the pair
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (both CTAs)
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            // shared::cluster address of the LEADER's full barriers (same offset, CTA rank 0)
            const uint32_t leader_full0 = mapa_shared(full_bar(0), 0);
            for (int t = cluster_id; t < num_tiles; t += num_clusters) {
                const int m_idx = t / p.num_n_tiles, n_idx = t % p.num_n_tiles;
                const int row_a = m_idx * 2 * C2_BM + rank * C2_BM;
                const int row_b = n_idx * C2_BN + rank * (C2_BN / 2);
                for (int kb = 0; kb < num_kb; ++kb) {
                    ptx::mbar_wait(empty_bar(stage), phase ^ 1);
                    const uint32_t sa = smem_base + stage * C2_STAGE_BYTES;
                    const uint32_t sb = sa + C2_A_BYTES;
                    if (leader) ptx::mbar_arrive_expect_tx(full_bar(stage), 2 * C2_STAGE_BYTES);
                    tma_load_2d_2sm(sa, &tmA, leader_full0 + 8u * stage, kb * C2_BK, row_a);
                    tma_load_2d_2sm(sb, &tmB, leader_full0 + 8u * stage, kb * C2_BK, row_b);
                    if (++stage == C2_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (leader CTA only)
        if (leader && lane == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(2 * C2_BM, C2_BN);
            int stage = 0, as = 0;
            uint32_t phase = 0, aphase = 0;
            for (int t = cluster_id; t < num_tiles; t += num_clusters) {
                ptx::mbar_wait(tempty_bar(as), aphase ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * C2_BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    ptx::mbar_wait(full_bar(stage), phase);
                    ptx::tc_fence_after();
                    const uint32_t sa = smem_base + stage * C2_STAGE_BYTES;
                    const uint64_t adesc = ptx::make_kmajor_sw128_desc(sa);
                    const uint64_t bdesc = ptx::make_kmajor_sw128_desc(sa + C2_A_BYTES);
#pragma unroll
                    for (int k = 0; k < C2_BK / 16; ++k)
                        umma_bf16_2cta(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    umma_commit_2cta(empty_bar(stage));   // frees the stage in BOTH CTAs
                    if (++stage == C2_STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit_2cta(tfull_bar(as));          // accumulator ready in BOTH CTAs
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue (8 warps per CTA, own 128 rows)
        const int q = warp & 3;
        const int col_half = (warp - 2) / 4;
        constexpr int NCHUNK = (ACT == EPI_ACT_SWIGLU) ? C2_BN / 64 : C2_BN / 32;
        constexpr int BN_OUT = (ACT == EPI_ACT_SWIGLU) ? C2_BN / 2 : C2_BN;
        constexpr int CH_PER = NCHUNK / (C2_EPI_WARPS / 4);
        const int chunk_lo = col_half * CH_PER, chunk_hi = chunk_lo + CH_PER;
        const int n_out_total = (ACT == EPI_ACT_SWIGLU) ? p.N / 2 : p.N;
        const uint32_t leader_tempty0 = mapa_shared(tempty_bar(0), 0);
        int as = 0;
        uint32_t aphase = 0;
        for (int t = cluster_id; t < num_tiles; t += num_clusters) {
            const int m_idx = t / p.num_n_tiles, n_idx = t % p.num_n_tiles;
            ptx::mbar_wait(tfull_bar(as), aphase);
            ptx::tc_fence_after();
            const int row = m_idx * 2 * C2_BM + rank * C2_BM + q * 32 + lane;
            const bool row_ok = row < p.M;
            const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + as * C2_BN;
            if constexpr (ACT == EPI_ACT_SWIGLU) {
#pragma unroll 1
                for (int c = chunk_lo; c < chunk_hi; ++c) {
                    uint32_t acc[32], accu[32];
                    ptx::tmem_ld_32x32(t_row + c * 32, acc);
                    ptx::tmem_ld_32x32(t_row + C2_BN / 2 + c * 32, accu);
                    ptx::tmem_wait_ld();
                    epilogue_chunk32<ACT, RES, OUT_F32>(p, acc, accu, row, row_ok, n_idx * C2_BN + c * 32,
                                                        n_idx * BN_OUT + c * 32, n_out_total);
                }
            } else {
                // software-pipelined: the tcgen05.ld of chunk i+1 is in flight while chunk i goes through the epilogue math and its
                // global stores (the single-buffered loop spent ~30 % of its stall samples waiting on the TMEM load)
                uint32_t acc[2][32];
                uint4 b4[2][4], r4[2][4];
                ptx::tmem_ld_32x32(t_row + chunk_lo * 32, acc[0]);
                epilogue_load_bias(p, n_idx * C2_BN + chunk_lo * 32, b4[0]);
                if (RES == EPI_RES_BF16) epilogue_load_res_bf16(p, row, row_ok, n_idx * BN_OUT + chunk_lo * 32, n_out_total, r4[0]);
#pragma unroll
                for (int i = 0; i < CH_PER; ++i) {
                    const int c = chunk_lo + i;
                    ptx::tmem_wait_ld();
                    if (i + 1 < CH_PER) {
                        ptx::tmem_ld_32x32(t_row + (c + 1) * 32, acc[(i + 1) & 1]);
                        epilogue_load_bias(p, n_idx * C2_BN + (c + 1) * 32, b4[(i + 1) & 1]);
                        if (RES == EPI_RES_BF16)
                            epilogue_load_res_bf16(p, row, row_ok, n_idx * BN_OUT + (c + 1) * 32, n_out_total, r4[(i + 1) & 1]);
                    }
                    epilogue_chunk32<ACT, RES, OUT_F32>(p, acc[i & 1], acc[i & 1], row, row_ok, n_idx * C2_BN + c * 32,
                                                        n_idx * BN_OUT + c * 32, n_out_total, b4[i & 1], r4[i & 1]);
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(leader_tempty0 + 8u * as);
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
    }

    ptx::tc_fence_before();
    cluster_sync_all();       // the peer may still be reading our smem / signalling our barriers until here
    if (warp == 1) {
        ptx::tc_fence_after();
        tmem_dealloc_2cta(tmem_base, C2_TMEM_COLS);
    }
}

template <int ACT, int RES, bool OUT_F32>
int launch_2cta(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams2& p, cudaStream_t stream) {
    auto kern = gemm_bf16_tcgen05_2cta_kernel<ACT, RES, OUT_F32>;
    static unsigned long long attr_devs = 0ull;
    if (first_use_on_device(attr_devs)) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C2_SMEM) != cudaSuccess) return GVL_ERR_CUDA;
    }
    int clusters = p.num_m_tiles * p.num_n_tiles;
    if (clusters > num_sms() / 2) clusters = num_sms() / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(C2_THREADS);
    cfg.dynamicSmemBytes = C2_SMEM;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, kern, tmA, tmB, p) != cudaSuccess) return GVL_ERR_CUDA;
    g_launch_count++;
    return cudaGetLastError() == cudaSuccess ? GVL_OK : GVL_ERR_CUDA;
}

}  // namespace

// Same contract as gemm_bf16 (gemm_tcgen05.cu); requires N tiles of 256 (SWIGLU: N % 256 == 0).
int gemm_bf16_2cta(const void* A, int lda, const void* W, int ldw, void* out, int ldo, int M, int N, int K,
                   const void* bias, const float* gamma, const void* residual, int ldr, int act, int res, int out_f32,
                   cudaStream_t stream) {
    CUtensorMap tmA, tmB;
    int rc = make_tmap_2d_bf16(&tmA, A, M, K, lda, C2_BM, C2_BK);
    if (rc != GVL_OK) return rc;
    rc = make_tmap_2d_bf16(&tmB, W, N, K, ldw, C2_BN / 2, C2_BK);
    if (rc != GVL_OK) return rc;
    GemmParams2 p;
    p.M = M; p.N = N; p.K = K;
    p.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
    p.gamma = gamma;
    p.residual = residual;
    p.out = out;
    p.ldo = ldo; p.ldr = ldr;
    p.num_m_tiles = (M + 2 * C2_BM - 1) / (2 * C2_BM);
    p.num_n_tiles = (N + C2_BN - 1) / C2_BN;
    if (act == EPI_ACT_NONE && res == EPI_RES_NONE && !out_f32) return launch_2cta<EPI_ACT_NONE, EPI_RES_NONE, false>(tmA, tmB, p, stream);
    if (act == EPI_ACT_GELU && res == EPI_RES_NONE && !out_f32) return launch_2cta<EPI_ACT_GELU, EPI_RES_NONE, false>(tmA, tmB, p, stream);
    if (act == EPI_ACT_QUICKGELU && res == EPI_RES_NONE && !out_f32) return launch_2cta<EPI_ACT_QUICKGELU, EPI_RES_NONE, false>(tmA, tmB, p, stream);
    if (act == EPI_ACT_SWIGLU && res == EPI_RES_NONE && !out_f32) return launch_2cta<EPI_ACT_SWIGLU, EPI_RES_NONE, false>(tmA, tmB, p, stream);
    if (act == EPI_ACT_NONE && res == EPI_RES_BF16 && !out_f32) return launch_2cta<EPI_ACT_NONE, EPI_RES_BF16, false>(tmA, tmB, p, stream);
    if (act == EPI_ACT_NONE && res == EPI_RES_F32 && out_f32) return launch_2cta<EPI_ACT_NONE, EPI_RES_F32, true>(tmA, tmB, p, stream);
    if (act == EPI_ACT_NONE && res == EPI_RES_NONE && out_f32) return launch_2cta<EPI_ACT_NONE, EPI_RES_NONE, true>(tmA, tmB, p, stream);
    return GVL_ERR_ARG;
}

}  // namespace gvl
