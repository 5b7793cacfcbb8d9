// Dense bf16 contraction C[M,N] = A[M,K] * W[N,K]^T on the sm_100a tensor cores.
//
// This single kernel family replaces every nn.Linear / patch-embed conv on the reference hot path
// (SURVEY 2.1 "Dense linears"): CLIP q/k/v/out/fc1/fc2 (modeling_clip.py:244-247, 336-337),
// InternVideo2 qkv/proj/fc1/fc2 (internvideo2.py:549-551, 624-627), projectors
// (llava_next_video.py:31-32, 46-47) and Phi-3 qkv/o/gate_up/down (modeling_phi3.py:453-454,
// 513-514), with the elementwise work that follows each of them fused into the epilogue.
//
// Structure (one CTA per SM, persistent over output tiles):
//   warp 0      TMA producer: cp.async.bulk.tensor (SWIZZLE_128B) A[128x64] + W[BNx64] per stage
//   warp 1      MMA issuer: one elected thread issues tcgen05.mma (M=128, N=BN, K=16), accumulator
//               in TMEM, double-buffered (2 x BN columns) so the epilogue overlaps the next tile
//   warps 2..5  epilogue: tcgen05.ld 32 lanes x 32 columns -> registers -> fused math -> global
//
// Rounding points mirror the reference bf16 autocast forward (SURVEY 8a "numerics contract"):
// the fp32 accumulator (+bias) is rounded to bf16 first (that is what nn.Linear returns), then
// activation / LayerScale / residual are applied with the same intermediate roundings.
#include <cuda.h>
#include <stdio.h>
#include "gvl_internal.h"
#include "ptx.cuh"

namespace gvl {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int GEMM_EPI_WARPS = 8;                       // two warps per TMEM lane quarter, each takes half of the columns
constexpr int GEMM_THREADS = 64 + 32 * GEMM_EPI_WARPS;

template <int BN>
struct GemmCfg {
    static constexpr int STAGES = (BN == 256) ? 4 : 6;
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
    static constexpr int TMEM_COLS = 2 * BN;
};

struct GemmParams {
    int M, N, K;
    const __nv_bfloat16* bias;  // [N] or nullptr
    const float* gamma;         // [N_out] LayerScale or nullptr
    const void* residual;       // [M, ldr] (bf16 if RES==1, f32 if RES==2)
    void* out;                  // [M, ldo] bf16 or f32
    int ldo, ldr;
    int num_m_tiles, num_n_tiles;
};

enum { ACT_NONE = 0, ACT_GELU = 1, ACT_QUICKGELU = 2, ACT_SWIGLU = 3 };
enum { RES_NONE = 0, RES_BF16 = 1, RES_F32 = 2 };

template <int BN, int ACT, int RES, bool OUT_F32>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA,
                         const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
    using Cfg = GemmCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B tiles need 1024-byte alignment.
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
    uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(smem_gen + STAGES * Cfg::STAGE_BYTES + 8 * (2 * STAGES + 4));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_kb = (p.K + BK - 1) / BK;
    const int num_tiles = p.num_m_tiles * p.num_n_tiles;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmA);
        ptx::prefetch_tmap(&tmB);
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(full_bar(s), 1);
            ptx::mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(tfull_bar(s), 1);
            ptx::mbar_init(tempty_bar(s), GEMM_EPI_WARPS);
        }
        ptx::fence_mbar_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                const int m_idx = t / p.num_n_tiles;
                const int n_idx = t % p.num_n_tiles;
                for (int kb = 0; kb < num_kb; ++kb) {
                    ptx::mbar_wait(empty_bar(stage), phase ^ 1);
                    const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
                    const uint32_t sb = sa + Cfg::A_BYTES;
                    ptx::mbar_arrive_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
                    ptx::tma_load_2d(sa, &tmA, full_bar(stage), kb * BK, m_idx * BM);
                    ptx::tma_load_2d(sb, &tmB, full_bar(stage), kb * BK, n_idx * BN);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                ptx::mbar_wait(tempty_bar(as), aphase ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    ptx::mbar_wait(full_bar(stage), phase);
                    ptx::tc_fence_after();
                    const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
                    const uint32_t sb = sa + Cfg::A_BYTES;
                    const uint64_t adesc = ptx::make_kmajor_sw128_desc(sa);
                    const uint64_t bdesc = ptx::make_kmajor_sw128_desc(sb);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        ptx::umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc,
                                       (kb > 0 || k > 0) ? 1u : 0u);
                    }
                    ptx::umma_commit(empty_bar(stage));  // smem slot free when these MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                ptx::umma_commit(tfull_bar(as));  // accumulator complete
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue (warps 2..5)
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        const int col_half = (warp - 2) / 4;  // warps 2..5 take the first half of the chunks, 6..9 the second
        int as = 0;
        uint32_t aphase = 0;
        constexpr int NCHUNK = (ACT == ACT_SWIGLU) ? BN / 64 : BN / 32;
        constexpr int BN_OUT = (ACT == ACT_SWIGLU) ? BN / 2 : BN;
        constexpr int CH_PER = NCHUNK / (GEMM_EPI_WARPS / 4);
        const int chunk_lo = col_half * CH_PER, chunk_hi = chunk_lo + CH_PER;
        const int n_out_total = (ACT == ACT_SWIGLU) ? p.N / 2 : p.N;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
            const int m_idx = t / p.num_n_tiles;
            const int n_idx = t % p.num_n_tiles;
            ptx::mbar_wait(tfull_bar(as), aphase);
            ptx::tc_fence_after();
            const int row = m_idx * BM + q * 32 + lane;
            const bool row_ok = row < p.M;
            const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + as * BN;
#pragma unroll 1
            for (int c = chunk_lo; c < chunk_hi; ++c) {
                uint32_t acc[32];
                float v[32];
                const int col_in = n_idx * BN + c * 32;       // column in the GEMM's N space
                const int col_out = n_idx * BN_OUT + c * 32;  // column in the output tensor
                ptx::tmem_ld_32x32(t_row + c * 32, acc);
                if (ACT == ACT_SWIGLU) {
                    uint32_t accu[32];
                    ptx::tmem_ld_32x32(t_row + BN / 2 + c * 32, accu);
                    ptx::tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        // Phi3MLP (modeling_phi3.py:458-464): up * silu(gate) on bf16 tensors.
                        float g = bf16r(__uint_as_float(acc[j]));
                        float u = bf16r(__uint_as_float(accu[j]));
                        v[j] = bf16r(u * bf16r(silu_f(g)));
                    }
                } else {
                    ptx::tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
                    if (p.bias != nullptr && col_in < p.N) {
                        const uint4* bp = reinterpret_cast<const uint4*>(p.bias + col_in);
#pragma unroll
                        for (int g8 = 0; g8 < 4; ++g8) {
                            if (col_in + g8 * 8 < p.N) {
                                uint4 b = __ldg(bp + g8);
                                float2 f;
                                f = unpack_bf16(b.x); v[g8 * 8 + 0] += f.x; v[g8 * 8 + 1] += f.y;
                                f = unpack_bf16(b.y); v[g8 * 8 + 2] += f.x; v[g8 * 8 + 3] += f.y;
                                f = unpack_bf16(b.z); v[g8 * 8 + 4] += f.x; v[g8 * 8 + 5] += f.y;
                                f = unpack_bf16(b.w); v[g8 * 8 + 6] += f.x; v[g8 * 8 + 7] += f.y;
                            }
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float x = bf16r(v[j]);  // what nn.Linear returns under bf16 autocast
                        if (ACT == ACT_GELU) x = bf16r(gelu_erf(x));
                        if (ACT == ACT_QUICKGELU) x = quick_gelu_bf16(x);
                        v[j] = x;
                    }
                }
                if (row_ok && col_out < n_out_total) {
                    if (p.gamma != nullptr) {
                        // LayerScale (internvideo2.py:451-466): fp32 multiply, rounded back to bf16.
#pragma unroll
                        for (int g4 = 0; g4 < 8; ++g4) {
                            if (col_out + g4 * 4 < n_out_total) {
                                float4 gm = __ldg(reinterpret_cast<const float4*>(p.gamma + col_out) + g4);
                                v[g4 * 4 + 0] = bf16r(v[g4 * 4 + 0] * gm.x);
                                v[g4 * 4 + 1] = bf16r(v[g4 * 4 + 1] * gm.y);
                                v[g4 * 4 + 2] = bf16r(v[g4 * 4 + 2] * gm.z);
                                v[g4 * 4 + 3] = bf16r(v[g4 * 4 + 3] * gm.w);
                            }
                        }
                    }
                    if (RES == RES_BF16) {
                        const uint4* rp = reinterpret_cast<const uint4*>(
                            reinterpret_cast<const __nv_bfloat16*>(p.residual) + size_t(row) * p.ldr + col_out);
#pragma unroll
                        for (int g8 = 0; g8 < 4; ++g8) {
                            if (col_out + g8 * 8 < n_out_total) {
                                uint4 r = rp[g8];
                                float2 f;
                                f = unpack_bf16(r.x); v[g8 * 8 + 0] += f.x; v[g8 * 8 + 1] += f.y;
                                f = unpack_bf16(r.y); v[g8 * 8 + 2] += f.x; v[g8 * 8 + 3] += f.y;
                                f = unpack_bf16(r.z); v[g8 * 8 + 4] += f.x; v[g8 * 8 + 5] += f.y;
                                f = unpack_bf16(r.w); v[g8 * 8 + 6] += f.x; v[g8 * 8 + 7] += f.y;
                            }
                        }
                    } else if (RES == RES_F32) {
                        const float4* rp = reinterpret_cast<const float4*>(
                            reinterpret_cast<const float*>(p.residual) + size_t(row) * p.ldr + col_out);
#pragma unroll
                        for (int g4 = 0; g4 < 8; ++g4) {
                            if (col_out + g4 * 4 < n_out_total) {
                                float4 r = rp[g4];
                                v[g4 * 4 + 0] += r.x; v[g4 * 4 + 1] += r.y;
                                v[g4 * 4 + 2] += r.z; v[g4 * 4 + 3] += r.w;
                            }
                        }
                    }
                    if (OUT_F32) {
                        float4* op = reinterpret_cast<float4*>(
                            reinterpret_cast<float*>(p.out) + size_t(row) * p.ldo + col_out);
#pragma unroll
                        for (int g4 = 0; g4 < 8; ++g4) {
                            if (col_out + g4 * 4 < n_out_total)
                                op[g4] = make_float4(v[g4 * 4], v[g4 * 4 + 1], v[g4 * 4 + 2], v[g4 * 4 + 3]);
                        }
                    } else {
                        uint4* op = reinterpret_cast<uint4*>(
                            reinterpret_cast<__nv_bfloat16*>(p.out) + size_t(row) * p.ldo + col_out);
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
            }
            // accumulator stage drained -> hand it back to the MMA warp
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(tempty_bar(as));
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) return nullptr;
        fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

// Row-major bf16 matrix [rows, cols] with leading dimension ld (elements); box = 64 cols x box_rows.
int make_tmap_2d_bf16(CUtensorMap* tm, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols) {
    PFN_encodeTiled fn = get_encode_fn();
    if (fn == nullptr) return GVL_ERR_DRIVER;
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld * 2) % 16 != 0) return GVL_ERR_ALIGN;
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {ld * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box,
                    estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? GVL_OK : GVL_ERR_DRIVER;
}

static int g_num_sms = 0;
int num_sms() {
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    return g_num_sms;
}

template <int BN, int ACT, int RES, bool OUT_F32>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p,
                       cudaStream_t stream) {
    using Cfg = GemmCfg<BN>;
    auto kern = gemm_bf16_tcgen05_kernel<BN, ACT, RES, OUT_F32>;
    static unsigned long long attr_devs = 0ull;
    if (first_use_on_device(attr_devs)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        if (e != cudaSuccess) return GVL_ERR_CUDA;
    }
    int grid = p.num_m_tiles * p.num_n_tiles;
    if (grid > num_sms()) grid = num_sms();
    kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, p);
    g_launch_count++;
    return cudaGetLastError() == cudaSuccess ? GVL_OK : GVL_ERR_CUDA;
}

template <int BN>
static int dispatch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, int act,
                         int res, int out_f32, cudaStream_t stream) {
    if (act == ACT_NONE && res == RES_NONE && !out_f32) return launch_gemm<BN, ACT_NONE, RES_NONE, false>(tmA, tmB, p, stream);
    if (act == ACT_GELU && res == RES_NONE && !out_f32) return launch_gemm<BN, ACT_GELU, RES_NONE, false>(tmA, tmB, p, stream);
    if (act == ACT_QUICKGELU && res == RES_NONE && !out_f32) return launch_gemm<BN, ACT_QUICKGELU, RES_NONE, false>(tmA, tmB, p, stream);
    if (act == ACT_SWIGLU && res == RES_NONE && !out_f32) return launch_gemm<BN, ACT_SWIGLU, RES_NONE, false>(tmA, tmB, p, stream);
    if (act == ACT_NONE && res == RES_BF16 && !out_f32) return launch_gemm<BN, ACT_NONE, RES_BF16, false>(tmA, tmB, p, stream);
    if (act == ACT_NONE && res == RES_F32 && out_f32) return launch_gemm<BN, ACT_NONE, RES_F32, true>(tmA, tmB, p, stream);
    if (act == ACT_NONE && res == RES_NONE && out_f32) return launch_gemm<BN, ACT_NONE, RES_NONE, true>(tmA, tmB, p, stream);
    return GVL_ERR_ARG;
}

// C = epilogue(A[M,K] @ W[N,K]^T).  See include/gvl.h: gvl_gemm_bf16.
int gemm_bf16(const void* A, int lda, const void* W, int ldw, void* out, int ldo, int M, int N, int K,
              const void* bias, const float* gamma, const void* residual, int ldr, int act, int res,
              int out_f32, int bn_hint, cudaStream_t stream) {
    if (M <= 0 || N <= 0 || K <= 0) return GVL_ERR_ARG;
    if (N % 8 != 0 || K % 8 != 0) return GVL_ERR_ARG;
    if (act == ACT_SWIGLU && (N % 256 != 0)) return GVL_ERR_ARG;
    int BN = bn_hint;
    if (BN != 128 && BN != 256) {
        // Prefer the 128x256 tile; fall back to 128x128 when N pads badly or the grid would be small.
        int waste256 = ((N + 255) / 256) * 256 - N;
        long tiles256 = long((M + BM - 1) / BM) * ((N + 255) / 256);
        BN = (waste256 * 8 > N || tiles256 < num_sms()) ? 128 : 256;
        if (act == ACT_SWIGLU) BN = 256;
    }
    // 2-CTA (cta_group::2) 256x256 tiles when the problem fills the 74 CTA pairs (bn_hint = 128 keeps the 1-CTA kernel: tests, A/B)
    {
        const long tiles2 = long((M + 255) / 256) * ((N + 255) / 256);
        if (BN == 256 && bn_hint != 128 && tiles2 >= num_sms() / 2) {
            prof_begin(GVL_PROF_GEMM, 2.0 * M * (double)N * K, stream);
            int rc2 = gemm_bf16_2cta(A, lda, W, ldw, out, ldo, M, N, K, bias, gamma, residual, ldr, act, res, out_f32, stream);
            prof_end(GVL_PROF_GEMM, stream);
            return rc2;
        }
    }
    CUtensorMap tmA, tmB;
    int rc = make_tmap_2d_bf16(&tmA, A, M, K, lda, BM, BK);
    if (rc != GVL_OK) return rc;
    rc = make_tmap_2d_bf16(&tmB, W, N, K, ldw, BN, BK);
    if (rc != GVL_OK) return rc;
    GemmParams p;
    p.M = M; p.N = N; p.K = K;
    p.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
    p.gamma = gamma;
    p.residual = residual;
    p.out = out;
    p.ldo = ldo; p.ldr = ldr;
    p.num_m_tiles = (M + BM - 1) / BM;
    p.num_n_tiles = (N + BN - 1) / BN;
    prof_begin(GVL_PROF_GEMM, 2.0 * M * (double)N * K, stream);
    rc = (BN == 256) ? dispatch_gemm<256>(tmA, tmB, p, act, res, out_f32, stream)
                     : dispatch_gemm<128>(tmA, tmB, p, act, res, out_f32, stream);
    prof_end(GVL_PROF_GEMM, stream);
    return rc;
}

}  // namespace gvl
