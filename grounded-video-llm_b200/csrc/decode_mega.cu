// One persistent kernel per decode step ("megakernel"): the 4 x L + 1 weight matrices of a step are streamed
// through a shared-memory ring by ONE producer warp per SM that never stops for phase boundaries (weights do not
// depend on activations), while 8 consumer warps per SM walk the phases of the step
//     per layer: [norm+qkv] B [rope + KV append + split-KV attention] B [merge + o_proj+res] B
//                [norm+gate_up+SwiGLU] B [down+res] B | [norm+lm_head+bias+argmax] B bookkeeping
// separated by grid-wide barriers B (atomic counter, all CTAs co-resident, cooperative launch). One launch runs ALL
// the steps of a gvl_lm_decode call (token, position and EOS state live in device memory).
//
// GEMV phases: a work item is 8 weight rows x <=512 k (8 KB). The decode path owns a PACKED copy of the weights
// (decode_mega_pack: item-major, inside an item [32-k chunk][row][32 k]) so that one item is ONE contiguous 8 KB
// cp.async.bulk into a consumer warp's ring slot -- tools/probe_stream.cu measured 4.8 TB/s with 1 KB row-wise copies
// and 7.2 TB/s with >= 4 KB copies on B200 -- and the interleave makes the fragment loads bank-conflict free without
// padding. The warp multiplies the item with mma.sync.m16n8k16 with the WEIGHTS as the A operand: the 16 bytes a lane reads
// from an item are its A quad (rows 0-7 = the 8 weight rows with the even bf16 pairs of each 8-k group, rows 8-15 the same rows
// with the odd pairs), the activation vector is staged pair-permuted as B columns 0 / 1 (gemv_items): one LDS.128 + one LDS.64 per
// HMMA, ~3 warp instructions per KB of weights, bounded by shared-memory bandwidth at ~3x the HBM rate. Inside a phase the
// consumers therefore wait for ring items (the work phases stream at the HBM copy peak) and the ring absorbs ~4 us of HBM stream
// while they sit in a barrier / stage the next activation vector. Items of a CTA are dealt round-robin to its 8 warps
// (k-split inside the CTA, partial sums reduced through shared memory in a fixed order -> deterministic).
// Every consumer warp has its own 3-slot ring fed by one lane of the producer warp, which keeps at most `inflight`
// copies per lane outstanding: 8 x 8 KB x 148 SMs = 9.5 MB in flight is enough for the full HBM rate, while deeper
// queues (the first version had 28 MB outstanding) only add queueing delay (4+ us measured) to every latency-critical
// load of the phase boundaries (activation staging, barrier atomics, q/k of the attention phase).
//
// Attention phase: every head gets the same number of consumer warps of the grid (37 at 148 SMs x 8 warps / 32 heads) and every
// warp one contiguous token range of one head (att_split); a warp streams K/V rows through its ring (4 lanes per token, 8 tokens
// per pass) with an online softmax, warps of a CTA merge through shared memory, and the <= G/H + 2 CTA partials per head are merged
// by every CTA while it stages the o_proj input (one L2 round trip, no atomics, no extra barrier).
//
// Code size matters: the consumer side is ONE loop over the phase descriptors, so every phase function is inlined once (6.4k SASS
// instructions for D = 96). Spelled out per layer it was 13.1k instructions = 210 KB, thrashed the instruction cache in exactly the
// latency-critical boundary code and cost 7 % of the step (profiles/r2_decode.md). Tracing lives in a separate TRACE instantiation.
//
// Reference semantics per phase: Phi3DecoderLayer / LlamaDecoderLayer with q_len = 1 (modeling_phi3.py:1034-1095,
// 629-775, 413-445; modeling_llama.py:699-760), lm_head + .float() (modeling_phi3.py:1525-1526), greedy pick of
// HF generate (llava_next_video.py:655-661). Rounding points identical to gemv.cu / decode.cu.
#include "gvl_internal.h"
#include "ptx.cuh"
#include "decode.h"
#include "decode_mega.h"
#include <stdlib.h>

namespace gvl {

namespace {

constexpr int MG_CONSUMERS = 8;
constexpr int MG_THREADS = 32 * (MG_CONSUMERS + 1);          // 8 consumer warps + the ring producer warp
constexpr int MG_SLOT_BYTES = MEGA_ROWS * MEGA_SEG * 2;          // 8192: one packed item
constexpr int MG_SLOTS = 3;                                      // ring depth per consumer warp
constexpr int MG_RING_BYTES = MG_CONSUMERS * MG_SLOTS * MG_SLOT_BYTES;   // 196608
constexpr int MG_SMEM_LIMIT = 232448 - 3072;                     // 227 KB opt-in minus static shared memory
constexpr int MG_ATT_SHORT = 128;                                // ctx <= this: one warp per head
// bulk copies each producer lane keeps outstanding: 1 while its phase is still ahead of the consumers (pure prefetch: tools/probe_exchange.cu
// measured 8-22 us per grid-wide exchange with 2-3 copies per lane queued against 4 us with one), all slots once the consumers wait for
// exactly these items (round-1 sweep of (ahead, current): (1,1) 2.50 (1,2) 2.57 (1,3) 2.52 (2,3) 2.47 ms / step)
constexpr int MG_INFLIGHT_AHEAD = 1, MG_INFLIGHT_CUR = MG_SLOTS;

// Split of the attention phase over the grid's consumer warps: every head gets the same number of warps (wph = 37 for 148 SMs and
// 32 heads) and every warp ONE contiguous token range of ONE head, <= lw tokens. The first version cut the flat (head, token) space
// evenly instead; a warp that straddled two heads then streamed two short segments = 8 ring items instead of 6, and with 3 slots per
// warp and ~5k cycles of bulk-copy latency under load an item costs a third of a round trip whatever its size: those 32 warps took
// 22k cycles against 12k and the whole grid waited for them at the barrier (profiles/r2_decode.md).
struct AttSplit {
    int wph, lw;        // warps per head, tokens per warp (multiple of 8)
};
__host__ __device__ inline AttSplit att_split(int ctx, int H, int G) {
    AttSplit a;
    if (ctx <= MG_ATT_SHORT) { a.wph = 1; a.lw = ctx; return a; }          // one warp per head
    a.wph = (G * MG_CONSUMERS) / H;
    a.lw = (((ctx + a.wph - 1) / a.wph) + 7) & ~7;
    return a;
}
// CTAs whose warps work on head h: [att_c0, att_c1]
__host__ __device__ inline int att_c0(int h, int wph) { return (h * wph) / MG_CONSUMERS; }
__host__ __device__ inline int att_c1(int h, int wph) { return (h * wph + wph - 1) / MG_CONSUMERS; }
__host__ __device__ inline int att_scratch_bytes(int D) { return MG_CONSUMERS * 2 * D * 4 + 2 * MG_CONSUMERS * (D + 4) * 4; }
// K / V rows of a (head, token range) are contiguous in the cache, so they travel through the same per-warp rings as the
// weights: a segment of `len` cached tokens is cut evenly into chunks of <= ct_max tokens (one ring slot), multiple of 8
__host__ __device__ inline int att_chunk_len(int len, int ct_max) {
    const int nch = (len + ct_max - 1) / ct_max;
    return (((len + nch - 1) / nch) + 7) & ~7;
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s_m(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {      // non-blocking
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ uint4 ldcg4(const void* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr) : "memory");
    return r;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
    uint2 r;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(addr) : "memory");
    return r;
}
// pairs of an 8-k group reordered for the B fragment of gemv_items: [P0 P1 P2 P3] -> [P0 P2 P1 P3]
__device__ __forceinline__ uint4 x_perm(uint4 v) { return make_uint4(v.x, v.z, v.y, v.w); }
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cbar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }   // the 256 consumer threads

// consumers-only grid barrier; `epoch` counts the barriers passed and is published for the producer warp
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& epoch, int tid, volatile unsigned* cta_epoch) {
    cbar();
    ++epoch;
    if (tid == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        const unsigned target = epoch * gridDim.x;
        if (ld_acquire_u32(counter) < target) {
            const long long t0 = clock64();
            while (ld_acquire_u32(counter) < target) {
                if (clock64() - t0 > 4000000000LL) {   // ~2 s: a CTA died or the grid is not co-resident
                    printf("gvl: decode_mega grid barrier timeout block %d epoch %u\n", blockIdx.x, epoch);
                    __trap();
                }
            }
        }
        __threadfence_block();
        *cta_epoch = epoch;                          // the producer warp gates its K / V stream on this (produce_attention)
    }
    cbar();
}

// optional per-CTA phase trace (clock64 at: x staged / work done / barrier passed), MegaPlan::trace != nullptr
struct Tracer {
    long long* p;
    __device__ __forceinline__ void mark(int tid) {
        if (p != nullptr && tid == 0) *p++ = clock64();
    }
};

__device__ __forceinline__ int row_base_of(const MegaOp& op, int unit, int sel) {
    const int o = unit * MEGA_ROWS;
    return op.act == 3 ? (o / 128) * 256 + (o % 128) + sel * 128 : o;
}

struct Smem {
    uint8_t* ring;      // [warp][slot][8 KB]
    uint8_t* xa;        // activation staging area / attention scratch
    float* part;        // per-item partial sums (8 floats per item)
    uint64_t* full;     // [warp][slot]
    uint64_t* empty;    // [warp][slot]
    float* red;
    float* rope;        // [2][128]: cos / sin row of the current position
};

struct NormPre {        // RMSNorm weights of the NEXT phase, loaded before the grid barrier (K <= 4096)
    uint4 v[2];
};
__device__ __forceinline__ void prefetch_norm(const MegaOp& op, NormPre& np, int tid) {
    if (op.norm_w == nullptr || op.K > 4096) return;
    const int kv = op.K / 8;
#pragma unroll
    for (int u = 0; u < 2; ++u)
        if (tid + u * 256 < kv) np.v[u] = __ldg(reinterpret_cast<const uint4*>(op.norm_w) + tid + u * 256);
}

__device__ __forceinline__ uint4 norm8(uint4 v, uint4 wv, float rstd) {
    uint4 o;
    float2 f, g;
    f = unpack_bf16(v.x); g = unpack_bf16(wv.x); o.x = pack_bf16(bf16r(f.x * rstd) * g.x, bf16r(f.y * rstd) * g.y);
    f = unpack_bf16(v.y); g = unpack_bf16(wv.y); o.y = pack_bf16(bf16r(f.x * rstd) * g.x, bf16r(f.y * rstd) * g.y);
    f = unpack_bf16(v.z); g = unpack_bf16(wv.z); o.z = pack_bf16(bf16r(f.x * rstd) * g.x, bf16r(f.y * rstd) * g.y);
    f = unpack_bf16(v.w); g = unpack_bf16(wv.w); o.w = pack_bf16(bf16r(f.x * rstd) * g.x, bf16r(f.y * rstd) * g.y);
    return o;
}
__device__ __forceinline__ float sumsq8(uint4 v) {
    float2 f;
    float ss;
    f = unpack_bf16(v.x); ss = f.x * f.x + f.y * f.y;
    f = unpack_bf16(v.y); ss += f.x * f.x + f.y * f.y;
    f = unpack_bf16(v.z); ss += f.x * f.x + f.y * f.y;
    f = unpack_bf16(v.w); ss += f.x * f.x + f.y * f.y;
    return ss;
}

// ------------------------------------------------------------------ x staging: vector in global memory (+ RMSNorm)
__device__ __forceinline__ void stage_x_vec(const MegaOp& op, const __nv_bfloat16* xsrc, const NormPre& np, const Smem& S,
                                            int tid, int warp, int lane, long long* sm = nullptr) {
    if (sm != nullptr && tid == 0) sm[0] = clock64();
    __nv_bfloat16* sx = reinterpret_cast<__nv_bfloat16*>(S.xa);
    const int kv = op.K / 8;
    if (op.norm_w == nullptr) {
        for (int i = tid; i < kv; i += 256) reinterpret_cast<uint4*>(sx)[i] = x_perm(ldcg4(xsrc + (size_t)i * 8));
        cbar();
        return;
    }
    if (op.K <= 4096) {
        // one round trip: x stays in registers across the block reduction, norm weights were prefetched
        uint4 xv[2];
        float ss = 0.f;
#pragma unroll
        for (int u = 0; u < 2; ++u)
            if (tid + u * 256 < kv) xv[u] = ldcg4(xsrc + (size_t)(tid + u * 256) * 8);   // written by other CTAs: L1-bypassing
#pragma unroll
        for (int u = 0; u < 2; ++u)
            if (tid + u * 256 < kv) ss += sumsq8(xv[u]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        if (lane == 0) S.red[warp] = ss;
        if (sm != nullptr && tid == 0) sm[1] = clock64();
        cbar();
        if (sm != nullptr && tid == 0) sm[2] = clock64();
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < MG_CONSUMERS; ++w) t += S.red[w];
        const float rstd = rsqrtf(t / op.K + op.eps);
#pragma unroll
        for (int u = 0; u < 2; ++u)
            if (tid + u * 256 < kv) reinterpret_cast<uint4*>(sx)[tid + u * 256] = x_perm(norm8(xv[u], np.v[u], rstd));
        cbar();
        if (sm != nullptr && tid == 0) sm[3] = clock64();
        return;
    }
    float ss = 0.f;
    for (int i = tid; i < kv; i += 256) {
        const uint4 v = ldcg4(xsrc + (size_t)i * 8);
        reinterpret_cast<uint4*>(sx)[i] = v;
        ss += sumsq8(v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) S.red[warp] = ss;
    cbar();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < MG_CONSUMERS; ++w) t += S.red[w];
    const float rstd = rsqrtf(t / op.K + op.eps);
    for (int i = tid; i < kv; i += 256)
        reinterpret_cast<uint4*>(sx)[i] = x_perm(norm8(reinterpret_cast<uint4*>(sx)[i], __ldg(reinterpret_cast<const uint4*>(op.norm_w) + i), rstd));
    cbar();
}

// ------------------------------------------------------------------ x staging: merge of the split-KV attention partials
// Every CTA needs the whole merged vector, so every CTA reads all H x (<= G/H + 2) partial rows (~75 KB at H = 32) from L2 right
// after a grid barrier, under the producer's HBM stream: what this step costs is L2 round trips, not bytes. The first version walked
// a thread's D/32 element groups one after the other (3 dependent round trips for Phi-3.5, 7.9k cycles per layer in the trace). Fast
// path (H <= 32, <= 6 partials per head): thread tid first issues the (max, sum) header of pair (head tid / 8, partial tid % 8) and
// ALL partial rows of its D/32 element groups, then the softmax weights are exchanged through shared memory: one round trip.
// Same arithmetic in the same order as the general path (bit-identical results).
template <int D>
__device__ __forceinline__ void stage_x_attn(const MegaPlan& P, const Smem& S, int ctx, int tid, long long* sm = nullptr) {
    __nv_bfloat16* sx = reinterpret_cast<__nv_bfloat16*>(S.xa);
    const int H = P.heads, D4 = D + 4, maxp = P.att_maxp;
    const int wph = att_split(ctx, H, gridDim.x).wph;
    if (sm != nullptr && tid == 0) sm[0] = clock64();
    constexpr int NS = 6, NG = D / 32;
    if (maxp <= NS && H <= 32) {
        float2* wl = reinterpret_cast<float2*>(S.part);          // [head][8] (softmax weight, partial sum) -- `part` is idle between GEMV phases
        const int h1 = tid >> 3, s1 = tid & 7;
        float2 hd1 = make_float2(-INFINITY, 0.f);
        if (h1 < H) {
            const int c0 = att_c0(h1, wph), c1 = att_c1(h1, wph);
            if (s1 <= c1 - c0) hd1 = __ldcg(reinterpret_cast<const float2*>(P.att_ws + ((size_t)h1 * maxp + s1) * D4));
        }
        float4 ov[NG][NS];
#pragma unroll
        for (int i = 0; i < NG; ++i) {
            const int e0 = (tid + i * 256) * 4, h = e0 / D, d0 = e0 - h * D;
            const bool live = h < H;
            int np = 0;
            if (live) {
                const int c0 = att_c0(h, wph), c1 = att_c1(h, wph);
                np = c1 - c0 + 1;
            }
            const float* base = P.att_ws + (size_t)h * maxp * D4 + 4 + d0;
#pragma unroll
            for (int s2 = 0; s2 < NS; ++s2)
                ov[i][s2] = s2 < np ? __ldcg(reinterpret_cast<const float4*>(base + (size_t)s2 * D4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float M = hd1.x;
        M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, 1));
        M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, 2));
        M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, 4));
        wl[tid] = make_float2(__expf(hd1.x - M), hd1.y);
        if (sm != nullptr && tid == 0) sm[1] = clock64();
        cbar();
        if (sm != nullptr && tid == 0) sm[2] = clock64();
#pragma unroll
        for (int i = 0; i < NG; ++i) {
            const int e0 = (tid + i * 256) * 4, h = e0 / D;
            if (h >= H) continue;
            float4 num = make_float4(0.f, 0.f, 0.f, 0.f);
            float den = 0.f;
#pragma unroll
            for (int s2 = 0; s2 < NS; ++s2) {
                const float2 w = wl[h * 8 + s2];
                den += w.x * w.y;
                num.x += w.x * ov[i][s2].x; num.y += w.x * ov[i][s2].y; num.z += w.x * ov[i][s2].z; num.w += w.x * ov[i][s2].w;
            }
            const float inv = den > 0.f ? 1.0f / den : 0.f;
            uint32_t* dst = reinterpret_cast<uint32_t*>(sx) + (e0 >> 3) * 4 + ((e0 >> 2) & 1);      // x_perm layout (general path below)
            dst[0] = pack_bf16(num.x * inv, num.y * inv);
            dst[2] = pack_bf16(num.z * inv, num.w * inv);
        }
        cbar();
        if (sm != nullptr && tid == 0) sm[3] = clock64();
        return;
    }
    for (int gi = tid; gi < H * D / 4; gi += 256) {
        const int e0 = gi * 4, h = e0 / D, d0 = e0 - h * D;
        const int c0 = att_c0(h, wph), c1 = att_c1(h, wph);
        const int np = c1 - c0 + 1;
        const float* base = P.att_ws + (size_t)h * maxp * D4;
        float4 num = make_float4(0.f, 0.f, 0.f, 0.f);
        float den = 0.f;
        {
            float M = -INFINITY;
            for (int s = 0; s < np; ++s) M = fmaxf(M, __ldcg(base + (size_t)s * D4));
#pragma unroll 4
            for (int s = 0; s < np; ++s) {
                const float2 hd = __ldcg(reinterpret_cast<const float2*>(base + (size_t)s * D4));
                const float4 ov = __ldcg(reinterpret_cast<const float4*>(base + (size_t)s * D4 + 4 + d0));
                const float w = __expf(hd.x - M);
                den += w * hd.y;
                num.x += w * ov.x; num.y += w * ov.y; num.z += w * ov.z; num.w += w * ov.w;
            }
        }
        const float inv = den > 0.f ? 1.0f / den : 0.f;
        // x_perm layout: elements e0..e0+3 are pairs (P0, P1) or (P2, P3) of their 8-k group -> words 0 / 2 or 1 / 3
        uint32_t* dst = reinterpret_cast<uint32_t*>(sx) + (e0 >> 3) * 4 + ((e0 >> 2) & 1);
        dst[0] = pack_bf16(num.x * inv, num.y * inv);
        dst[2] = pack_bf16(num.z * inv, num.w * inv);
    }
    cbar();
}

// ------------------------------------------------------------------ consumer side of one GEMV phase (x already staged)
__device__ __forceinline__ void gemv_items(const MegaOp& op, const MegaPlan& P, const __nv_bfloat16* emb_row,
                                           float* extra_out, const Smem& S, uint32_t& cnt, int tid, int warp, int lane,
                                           long long* occ = nullptr) {
    const int G = gridDim.x, c = blockIdx.x;
    const int nu = op.units > c ? (op.units - c + G - 1) / G : 0;
    const int nsel = op.act == 3 ? 2 : 1;
    const int nseg = op.nseg;
    const int ipu = nsel * nseg;
    const int n_items = nu * ipu;
    const int g = lane >> 2, t = lane & 3;
    // residual of this thread's output column: loaded now, used in the epilogue (hides one L2 round trip)
    const __nv_bfloat16* res = (op.from_embed & 2) ? emb_row : op.residual;
    auto load_res = [&](int n) -> float {
        return __uint_as_float((uint32_t)__ldcg(reinterpret_cast<const unsigned short*>(res) + n) << 16);
    };
    if (occ != nullptr && tid == 0) occ[128] = clock64();        // tracing: entry / item loop done / CTA barrier passed / epilogue done
    float res_pre = 0.f;
    if (res != nullptr && tid < nu * MEGA_ROWS) {
        const int n = (c + (tid >> 3) * G) * MEGA_ROWS + (tid & 7);
        if (n < op.n_out) res_pre = load_res(n);
    }
    if (occ != nullptr && lane == 0) {
        // bring-up: how many of this warp's next slots have already landed when the phase starts (prefetch depth)
        int ready = 0;
        for (uint32_t k = 0; k < (uint32_t)MG_SLOTS; ++k) {
            const uint32_t cc = cnt + k;
            ready += mbar_test(ptx::smem_u32(&S.full[warp * MG_SLOTS + cc % MG_SLOTS]), (cc / MG_SLOTS) & 1) ? 1 : 0;
        }
        occ[warp] = ready;
    }
    // Fragment mapping (mma.m16n8k16, weights as the A operand): the 16 bytes thread (g, t) reads from a ring item are 8 consecutive k
    // of weight row g = four bf16 pairs P0..P3, used as a0..a3 = (row g, k-slot 0), (row g+8, slot 0), (row g, slot 1), (row g+8, slot 1):
    // the upper 8 rows of the tile are weight row g again, with the OTHER half of the k values. The activation vector is staged with
    // the pairs of each 8-k group reordered to [P0 P2 P1 P3] (x_perm), so column 0 of B (lanes g even) holds x for (P0, P2) and column 1
    // (lanes g odd) x for (P1, P3): C[g][0] + C[g+8][1] is the dot product of row g over the 32-k chunk. One LDS.128 + one LDS.64
    // per HMMA, no register shuffling, 512 weight bytes per HMMA (the first version used the weights as B: 256 bytes per HMMA plus
    // four MOVs to duplicate x into both row halves, and the loop was HMMA-issue bound, profiles/r2_decode.md).
    const uint32_t ring_w = ptx::smem_u32(S.ring) + warp * (MG_SLOTS * MG_SLOT_BYTES) + g * 64 + t * 16;   // item: [chunk][row g][64 B]
    const uint32_t x_u32 = ptx::smem_u32(S.xa) + t * 16 + (g & 1) * 8;
    const uint32_t full_u32 = ptx::smem_u32(S.full) + warp * (MG_SLOTS * 8), empty_u32 = ptx::smem_u32(S.empty) + warp * (MG_SLOTS * 8);
    // The per-item partial sums of a CTA's units usually fit the `part` buffer at once (one pass). A phase whose item count exceeds it
    // (Llama-3's 128558-row lm_head: 109 units x 8 segments per CTA) runs in passes of `upp` units; upp * ipu is a multiple of 8, so
    // item q still belongs to warp q % 8, which is the order the producer lane of that warp streams them in.
    int upp = nu;
    if (n_items > P.part_items) {
        const int mult = 8 / ((ipu & 1) ? 1 : (ipu & 2) ? 2 : (ipu & 4) ? 4 : 8);      // 8 / gcd(ipu, 8)
        upp = (P.part_items / ipu) / mult * mult;
    }
    unsigned long long key = 0ull;
    long long wait_clk = 0;                                  // tracing only: cycles this warp spent waiting for ring items, items consumed
    int n_waited = 0;
    // ring position and k segment are carried incrementally (no division per item): item q covers segment q % nseg
    uint32_t slot = cnt % MG_SLOTS, par = (cnt / MG_SLOTS) & 1;
    const int seg_bytes = op.seg_len * 2, seg_step = MG_CONSUMERS % nseg;
    const int nquad = op.seg_len / 128, tail2 = (op.seg_len / 64) & 1;     // chunks of 32 k: 4 per quad + one optional pair
    float* const part_w = S.part + g;
    for (int u0 = 0; u0 < nu; u0 += upp) {
        const int nu_p = min(upp, nu - u0);
        const int q_lo = u0 * ipu, q_hi = (u0 + nu_p) * ipu;
        int seg = (q_lo + warp) % nseg;
        for (int q = q_lo + warp; q < q_hi; q += MG_CONSUMERS) {
            const long long w0 = occ ? clock64() : 0;
            ptx::mbar_wait(full_u32 + slot * 8, par);
            if (occ) { wait_clk += clock64() - w0; ++n_waited; }
            uint32_t wa = ring_w + slot * MG_SLOT_BYTES;
            uint32_t xa = x_u32 + seg * seg_bytes;
            float acc[4][4];
#pragma unroll
            for (int a = 0; a < 4; ++a) acc[a][0] = acc[a][1] = acc[a][2] = acc[a][3] = 0.f;
#pragma unroll 2
            for (int ch = 0; ch < nquad; ++ch) {
                const uint4 w0 = lds128(wa), w1 = lds128(wa + 512), w2 = lds128(wa + 1024), w3 = lds128(wa + 1536);
                const uint2 x0 = lds64(xa), x1 = lds64(xa + 64), x2 = lds64(xa + 128), x3 = lds64(xa + 192);
                mma16816(acc[0], w0.x, w0.y, w0.z, w0.w, x0.x, x0.y);
                mma16816(acc[1], w1.x, w1.y, w1.z, w1.w, x1.x, x1.y);
                mma16816(acc[2], w2.x, w2.y, w2.z, w2.w, x2.x, x2.y);
                mma16816(acc[3], w3.x, w3.y, w3.z, w3.w, x3.x, x3.y);
                wa += 2048;
                xa += 256;
            }
            if (tail2) {
                const uint4 w0 = lds128(wa), w1 = lds128(wa + 512);
                const uint2 x0 = lds64(xa), x1 = lds64(xa + 64);
                mma16816(acc[0], w0.x, w0.y, w0.z, w0.w, x0.x, x0.y);
                mma16816(acc[1], w1.x, w1.y, w1.z, w1.w, x1.x, x1.y);
            }
            if (t == 0)
                part_w[(q - q_lo) * 8] = ((acc[0][0] + acc[0][3]) + (acc[1][0] + acc[1][3])) + ((acc[2][0] + acc[2][3]) + (acc[3][0] + acc[3][3]));
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(empty_u32 + slot * 8);
            ++cnt;
            if (++slot == MG_SLOTS) { slot = 0; par ^= 1; }
            seg += seg_step;
            if (seg >= nseg) seg -= nseg;
        }
        if (occ != nullptr && tid == 0) occ[129] = clock64();
        cbar();
        if (occ != nullptr && tid == 0) occ[130] = clock64();
        // ---- epilogue of this pass: one thread per output column, segments summed in a fixed order
        for (int o0 = 0; o0 < nu_p * MEGA_ROWS; o0 += 256) {
            const int o = o0 + tid;
            const bool active = o < nu_p * MEGA_ROWS;
            const int jl = o >> 3, col = o & 7;
            const int n = (c + (u0 + jl) * G) * MEGA_ROWS + col;
            const bool live = active && n < op.n_out;
            if (live) {
                const float* pp = S.part + (size_t)jl * ipu * 8 + col;
                float a0 = 0.f;
                for (int s = 0; s < nseg; ++s) a0 += pp[s * 8];
                if (op.act == 3) {
                    float a1 = 0.f;
                    for (int s = 0; s < nseg; ++s) a1 += pp[(nseg + s) * 8];
                    const float gt = bf16r(a0), u = bf16r(a1);
                    reinterpret_cast<__nv_bfloat16*>(op.out)[n] = __float2bfloat16_rn(u * bf16r(silu_f(gt)));
                } else {
                    float y = a0;
                    if (op.bias) y += __bfloat162float(op.bias[n]);
                    y = bf16r(y);
                    if (res) {
                        const float rv = (u0 == 0 && o < 256) ? res_pre : load_res(n);
                        y = bf16r(y + rv);
                    }
                    if (op.out_f32) reinterpret_cast<float*>(op.out)[n] = y;
                    else reinterpret_cast<__nv_bfloat16*>(op.out)[n] = __float2bfloat16_rn(y);
                    if (extra_out) extra_out[n] = y;
                    if (op.argmax) {
                        uint32_t u = __float_as_uint(y);
                        u ^= (u >> 31) ? 0xffffffffu : 0x80000000u;
                        const unsigned long long k = ((unsigned long long)u << 32) | (unsigned long long)(0xffffffffu - (uint32_t)n);
                        key = k > key ? k : key;
                    }
                }
            }
        }
        if (u0 + upp < nu) cbar();                       // the next pass overwrites the partial sums
    }
    if (occ != nullptr && lane == 0) { occ[32 + warp] = wait_clk; occ[64 + warp] = n_waited; }
    if (occ != nullptr && tid == 0) occ[131] = clock64();
    if (op.argmax) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
            key = other > key ? other : key;
        }
        if (lane == 0 && key != 0ull) atomicMax(P.amax, key);
    }
}

// ------------------------------------------------------------------ producer side of one GEMV phase: lane w feeds warp w
// `inflight` copies per lane while the phase is still ahead of the consumers (pure prefetch: deep queues only delay the
// latency-critical loads of the phase boundary the consumers are in), `inflight_cur` once the consumers have reached
// this phase (cta_epoch >= my_epoch) and are waiting for exactly these items.
__device__ __forceinline__ void produce_phase(const MegaOp& op, const Smem& S, uint32_t& pc, int inflight, int inflight_cur,
                                              volatile unsigned* cta_epoch, unsigned my_epoch, int lane) {
    const int G = gridDim.x, c = blockIdx.x;
    const int nu = op.units > c ? (op.units - c + G - 1) / G : 0;
    const int ipu = (op.act == 3 ? 2 : 1) * op.nseg;
    const int n_items = nu * ipu;
    const uint32_t item_bytes = (uint32_t)op.seg_len * 2 * MEGA_ROWS;
    const int my_n = (lane < MG_CONSUMERS && n_items > lane) ? (n_items - lane + MG_CONSUMERS - 1) / MG_CONSUMERS : 0;
    const uint32_t ring_w = ptx::smem_u32(S.ring) + lane * (MG_SLOTS * MG_SLOT_BYTES);
    const uint32_t full0 = ptx::smem_u32(S.full + lane * MG_SLOTS), empty0 = ptx::smem_u32(S.empty + lane * MG_SLOTS);
    int m = 0;
    while (__any_sync(0xffffffffu, m < my_n)) {
        bool issued = false;
        if (m < my_n) {
            const uint32_t slot = pc % MG_SLOTS, par = (pc / MG_SLOTS) & 1;
            bool ok = mbar_test(empty0 + slot * 8, par ^ 1);                    // slot consumed
            const uint32_t inf = *cta_epoch >= my_epoch ? inflight_cur : inflight;
            if (ok && pc >= inf) {                                              // bounded number of copies in flight
                const uint32_t pp = pc - inf;
                ok = mbar_test(full0 + (pp % MG_SLOTS) * 8, (pp / MG_SLOTS) & 1);
            }
            if (ok) {
                const int q = lane + m * MG_CONSUMERS;
                const int j = q / ipu, r = q - j * ipu;
                const size_t item = (size_t)(c + j * G) * ipu + r;              // packed order: [unit][sel][seg]
                ptx::mbar_arrive_expect_tx(full0 + slot * 8, item_bytes);
                bulk_g2s_m(ring_w + slot * MG_SLOT_BYTES, reinterpret_cast<const uint8_t*>(op.W) + item * item_bytes, item_bytes,
                           full0 + slot * 8);
                ++pc;
                ++m;
                issued = true;
            }
        }
        if (!__any_sync(0xffffffffu, issued)) __nanosleep(40);
    }
}

// ------------------------------------------------------------------ producer side of the attention phase: lane w feeds warp w
// the K / V chunks of its (head, token) range, in exactly the order attention_phase consumes them
template <int D>
__device__ __forceinline__ void produce_attention(const MegaPlan& P, int layer, int pos, const Smem& S, uint32_t& pc, int inflight,
                                                  int inflight_cur, unsigned my_epoch, int lane, volatile unsigned* cta_epoch,
                                                  unsigned need_epoch) {
    constexpr int ROWB = D * 2, CTM = (MG_SLOT_BYTES / ROWB) & ~7;
    const int H = P.heads, KVH = P.kv_heads, rep = H / KVH;
    const int G = gridDim.x, c = blockIdx.x;
    const int ctx = pos + 1;
    const AttSplit sp = att_split(ctx, H, G);
    const __nv_bfloat16* kc_l = P.kv + (size_t)layer * 2 * KVH * P.max_ctx * D;
    const __nv_bfloat16* vc_l = kc_l + (size_t)KVH * P.max_ctx * D;
    int tb = 0, tend = 0, cs = 8, hk = 0, is_v = 0;
    bool active = false;
    if (lane < MG_CONSUMERS) {
        const int gw = c * MG_CONSUMERS + lane, h = gw / sp.wph;
        const int t0 = (gw - h * sp.wph) * sp.lw;
        const int te = min(min(ctx, t0 + sp.lw), pos);       // the new token (pos) never comes from the cache
        if (h < H && te > t0) {
            tb = t0; tend = te; hk = h / rep;
            cs = att_chunk_len(te - t0, CTM);
            active = true;
        }
    }
    const uint32_t ring_w = ptx::smem_u32(S.ring) + lane * (MG_SLOTS * MG_SLOT_BYTES);
    const uint32_t full0 = ptx::smem_u32(S.full + lane * MG_SLOTS), empty0 = ptx::smem_u32(S.empty + lane * MG_SLOTS);
    // Rows appended by the previous step of this launch are part of this stream: wait until this CTA's consumers are past
    // the grid barrier that followed that step's attention phase of this layer. A warp without GEMV items in between
    // (tiny models) would otherwise let its ring run ahead of the append; at production sizes the wait never spins.
    while (*cta_epoch < need_epoch) __nanosleep(100);
    __threadfence_block();
    fence_proxy_async_global();
    while (__any_sync(0xffffffffu, active)) {
        bool issued = false;
        if (active) {
            const uint32_t slot = pc % MG_SLOTS, par = (pc / MG_SLOTS) & 1;
            bool ok = mbar_test(empty0 + slot * 8, par ^ 1);
            const uint32_t inf = *cta_epoch >= my_epoch ? inflight_cur : inflight;
            if (ok && pc >= inf) {
                const uint32_t pp = pc - inf;
                ok = mbar_test(full0 + (pp % MG_SLOTS) * 8, (pp / MG_SLOTS) & 1);
            }
            if (ok) {
                const int n = min(cs, tend - tb);
                const uint32_t bytes = (uint32_t)n * ROWB;
                const __nv_bfloat16* src = (is_v ? vc_l : kc_l) + ((size_t)hk * P.max_ctx + tb) * D;
                ptx::mbar_arrive_expect_tx(full0 + slot * 8, bytes);
                bulk_g2s_m(ring_w + slot * MG_SLOT_BYTES, src, bytes, full0 + slot * 8);
                ++pc;
                issued = true;
                if (is_v) {
                    tb += cs;
                    is_v = 0;
                    if (tb >= tend) active = false;
                } else {
                    is_v = 1;
                }
            }
        }
        if (!__any_sync(0xffffffffu, issued)) __nanosleep(40);
    }
}

// ------------------------------------------------------------------ attention phase
template <int D, bool TRACE>
__device__ __forceinline__ void attention_phase(const MegaPlan& P, int layer, int pos, const Smem& S, uint32_t& cnt, int warp,
                                                int lane) {
    constexpr int EPL = D / 4, VPL = EPL / 8, half = D / 2, D4 = D + 4;
    constexpr int ROWB = D * 2, CTM = (MG_SLOT_BYTES / ROWB) & ~7, NPASS = CTM / 8;   // tokens / 8-token passes per ring item
    const int H = P.heads, KVH = P.kv_heads, rep = H / KVH;
    const int G = gridDim.x, c = blockIdx.x;
    const int ctx = pos + 1;
    const AttSplit sp = att_split(ctx, H, G);
    const __nv_bfloat16* qkv = P.qkv;
    __nv_bfloat16* kc_l = P.kv + (size_t)layer * 2 * KVH * P.max_ctx * D;
    __nv_bfloat16* vc_l = kc_l + (size_t)KVH * P.max_ctx * D;
    const float* rcos = S.rope;
    const float* rsin = S.rope + 128;
    float* sq = reinterpret_cast<float*>(S.xa) + warp * 2 * D;     // rotated q [D]
    float* sk = sq + D;                                            // rotated new k [D]
    float* segs = reinterpret_cast<float*>(S.xa) + MG_CONSUMERS * 2 * D;   // [16][D + 4]: head, m, l, -, o[D]
    const int sub = lane & 3, tg = lane >> 2;          // 4 lanes per token, 8 tokens per pass
    const float scale = P.scale;

    // bring-up trace of the LAST layer: per-warp clock64 at 8 points of this phase (MEGA_TRACE_ATT_OFF + warp * 8 + k)
    long long* atr = (TRACE && layer == P.n_layers - 1 && lane == 0)
                         ? P.trace + (size_t)blockIdx.x * MEGA_TRACE_STRIDE + MEGA_TRACE_ATT_OFF + warp * 8 : nullptr;
#define ATR(k) do { if (atr) atr[k] = clock64(); } while (0)
    ATR(0);
    if (lane < 2) segs[(warp * 2 + lane) * D4] = __int_as_float(-1);
    const int gw = c * MG_CONSUMERS + warp;
    const int h = gw / sp.wph, hk = h / rep;
    const int t0 = (gw - h * sp.wph) * sp.lw;
    const int t1 = min(ctx, t0 + sp.lw);
    int sg = 0;
    if (h < H && t0 < t1) {
        // ---- RoPE of q (and of the new k) for this head: q_embed = bf16(bf16(q*cos) + bf16(rot(q)*sin))
        __syncwarp();
        {
            const __nv_bfloat16* qh = qkv + (size_t)h * D;
            const __nv_bfloat16* kh = qkv + (size_t)(H + hk) * D;
            float q1[(half + 31) / 32], q2[(half + 31) / 32], k1[(half + 31) / 32], k2[(half + 31) / 32];
#pragma unroll
            for (int u = 0; u < (half + 31) / 32; ++u) {
                const int j = lane + u * 32;
                if (j < half) {
                    q1[u] = __bfloat162float(__ldcg(qh + j)); q2[u] = __bfloat162float(__ldcg(qh + j + half));
                    k1[u] = __bfloat162float(__ldcg(kh + j)); k2[u] = __bfloat162float(__ldcg(kh + j + half));
                }
            }
#pragma unroll
            for (int u = 0; u < (half + 31) / 32; ++u) {
                const int j = lane + u * 32;
                if (j < half) {
                    const float c1 = rcos[j], s1 = rsin[j], c2 = rcos[j + half], s2 = rsin[j + half];
                    sq[j] = bf16r(bf16r(q1[u] * c1) + bf16r(-q2[u] * s1));
                    sq[j + half] = bf16r(bf16r(q2[u] * c2) + bf16r(q1[u] * s2));
                    sk[j] = bf16r(bf16r(k1[u] * c1) + bf16r(-k2[u] * s1));
                    sk[j + half] = bf16r(bf16r(k2[u] * c2) + bf16r(k1[u] * s2));
                }
            }
        }
        __syncwarp();
        ATR(1);
        const bool has_new = pos >= t0 && pos < t1;    // pos is the last token: has_new => t1 == ctx
        const int tend = has_new ? pos : t1;
        const __nv_bfloat16* vnew = qkv + (size_t)(H + KVH + hk) * D;
        uint4 vn[VPL];                                  // this lane's slice of the new v row
        if (has_new) {
#pragma unroll
            for (int i = 0; i < VPL; ++i)
                vn[i] = ldcg4(vnew + sub * EPL + i * 8);
            if ((h % rep) == 0 && lane < D / 8) {
                // exactly one segment per kv head appends the new row to the cache (for FUTURE steps; this step uses the
                // locally rotated copy, so there is no intra-phase dependency on this write)
                uint4 kq;
                const float* s8 = sk + lane * 8;
                kq.x = pack_bf16(s8[0], s8[1]); kq.y = pack_bf16(s8[2], s8[3]);
                kq.z = pack_bf16(s8[4], s8[5]); kq.w = pack_bf16(s8[6], s8[7]);
                *reinterpret_cast<uint4*>(kc_l + ((size_t)hk * P.max_ctx + pos) * D + lane * 8) = kq;
                *reinterpret_cast<uint4*>(vc_l + ((size_t)hk * P.max_ctx + pos) * D + lane * 8) = ldcg4(vnew + lane * 8);
                fence_proxy_async_global();        // later steps read this row with bulk copies (async proxy)
            }
        }
        ATR(2);
        float qf[EPL];
#pragma unroll
        for (int i = 0; i < EPL; ++i) qf[i] = sq[sub * EPL + i];
        float o[EPL];
#pragma unroll
        for (int i = 0; i < EPL; ++i) o[i] = 0.f;
        float m_run = -INFINITY, l_run = 0.f;
        // cached tokens [t0, tend): K chunk, then V chunk of the same tokens, from this warp's ring (produce_attention
        // issues the identical sequence). Row r of an item sits at r * ROWB; lane (tg, sub) reads 16-byte pieces at
        // tg * ROWB + sub * EPL * 2 + 16 i: conflict-free for D = 96 (8-lane phases hit 8 distinct 16-byte bank groups).
        if (t0 < tend) {
            const int cs = att_chunk_len(tend - t0, CTM);
            const uint32_t ring_w = ptx::smem_u32(S.ring) + warp * (MG_SLOTS * MG_SLOT_BYTES) + sub * (EPL * 2);
            for (int tb = t0; tb < tend; tb += cs) {
                const int n = min(cs, tend - tb);
                uint32_t slot = cnt % MG_SLOTS, par = (cnt / MG_SLOTS) & 1;
                ptx::mbar_wait(ptx::smem_u32(&S.full[warp * MG_SLOTS + slot]), par);
                uint32_t base = ring_w + slot * MG_SLOT_BYTES;
                float sc[NPASS];
                float bmax = -INFINITY;
#pragma unroll
                for (int pp = 0; pp < NPASS; ++pp) {
                    sc[pp] = -INFINITY;
                    if (pp * 8 < n) {
                        const int r = pp * 8 + tg;
                        const bool valid = r < n;
                        const uint32_t a = base + (valid ? r : 0) * ROWB;
                        float sv = 0.f;
#pragma unroll
                        for (int i = 0; i < VPL; ++i) {
                            const uint4 v = lds128(a + i * 16);
                            float2 f2;
                            f2 = unpack_bf16(v.x); sv += qf[i * 8 + 0] * f2.x + qf[i * 8 + 1] * f2.y;
                            f2 = unpack_bf16(v.y); sv += qf[i * 8 + 2] * f2.x + qf[i * 8 + 3] * f2.y;
                            f2 = unpack_bf16(v.z); sv += qf[i * 8 + 4] * f2.x + qf[i * 8 + 5] * f2.y;
                            f2 = unpack_bf16(v.w); sv += qf[i * 8 + 6] * f2.x + qf[i * 8 + 7] * f2.y;
                        }
                        sv += __shfl_xor_sync(0xffffffffu, sv, 1);
                        sv += __shfl_xor_sync(0xffffffffu, sv, 2);
                        sv = valid ? sv * scale : -INFINITY;
                        sc[pp] = sv;
                        bmax = fmaxf(bmax, sv);
                    }
                }
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&S.empty[warp * MG_SLOTS + slot]));
                ++cnt;
#pragma unroll
                for (int off = 4; off < 32; off <<= 1) bmax = fmaxf(bmax, __shfl_xor_sync(0xffffffffu, bmax, off));
                const float m_new = fmaxf(m_run, bmax);            // finite: n >= 1
                const float corr = __expf(m_run - m_new);
                l_run *= corr;
#pragma unroll
                for (int i = 0; i < EPL; ++i) o[i] *= corr;
                slot = cnt % MG_SLOTS; par = (cnt / MG_SLOTS) & 1;
                ptx::mbar_wait(ptx::smem_u32(&S.full[warp * MG_SLOTS + slot]), par);
                base = ring_w + slot * MG_SLOT_BYTES;
#pragma unroll
                for (int pp = 0; pp < NPASS; ++pp) {
                    if (pp * 8 < n) {
                        const int r = pp * 8 + tg;
                        const uint32_t a = base + (r < n ? r : 0) * ROWB;
                        const float p = bf16r(__expf(sc[pp] - m_new));      // 0 for the padding rows (sc = -inf)
                        l_run += p;
#pragma unroll
                        for (int i = 0; i < VPL; ++i) {
                            const uint4 v = lds128(a + i * 16);
                            float2 f2;
                            f2 = unpack_bf16(v.x); o[i * 8 + 0] += p * f2.x; o[i * 8 + 1] += p * f2.y;
                            f2 = unpack_bf16(v.y); o[i * 8 + 2] += p * f2.x; o[i * 8 + 3] += p * f2.y;
                            f2 = unpack_bf16(v.z); o[i * 8 + 4] += p * f2.x; o[i * 8 + 5] += p * f2.y;
                            f2 = unpack_bf16(v.w); o[i * 8 + 6] += p * f2.x; o[i * 8 + 7] += p * f2.y;
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&S.empty[warp * MG_SLOTS + slot]));
                ++cnt;
                m_run = m_new;
            }
        }
        ATR(3);
        if (has_new) {
            float sv = 0.f;
#pragma unroll
            for (int i = 0; i < EPL; ++i) sv += qf[i] * sk[sub * EPL + i];
            sv += __shfl_xor_sync(0xffffffffu, sv, 1);
            sv += __shfl_xor_sync(0xffffffffu, sv, 2);
            sv *= scale;
            const float m_new = fmaxf(m_run, sv);
            const float corr = __expf(m_run - m_new);
            l_run *= corr;
#pragma unroll
            for (int i = 0; i < EPL; ++i) o[i] *= corr;
            if (tg == 0) {
                const float p = bf16r(__expf(sv - m_new));
                l_run += p;
#pragma unroll
                for (int i = 0; i < VPL; ++i) {
                    const uint4 v = vn[i];
                    float2 f2;
                    f2 = unpack_bf16(v.x); o[i * 8 + 0] += p * f2.x; o[i * 8 + 1] += p * f2.y;
                    f2 = unpack_bf16(v.y); o[i * 8 + 2] += p * f2.x; o[i * 8 + 3] += p * f2.y;
                    f2 = unpack_bf16(v.z); o[i * 8 + 4] += p * f2.x; o[i * 8 + 5] += p * f2.y;
                    f2 = unpack_bf16(v.w); o[i * 8 + 6] += p * f2.x; o[i * 8 + 7] += p * f2.y;
                }
            }
            m_run = m_new;
        }
        ATR(4);
        // ---- reduce over the 8 token groups; the 4 sub-lanes of a token carry identical p, so l is reduced over tg only
#pragma unroll
        for (int off = 4; off < 32; off <<= 1) l_run += __shfl_xor_sync(0xffffffffu, l_run, off);
#pragma unroll
        for (int i = 0; i < EPL; ++i) {
            o[i] += __shfl_xor_sync(0xffffffffu, o[i], 4);
            o[i] += __shfl_xor_sync(0xffffffffu, o[i], 8);
            o[i] += __shfl_xor_sync(0xffffffffu, o[i], 16);
        }
        float* rec = segs + (size_t)(warp * 2 + sg) * D4;
        if (lane < 4) {
#pragma unroll
            for (int i = 0; i < EPL; ++i) rec[4 + lane * EPL + i] = o[i];
        }
        if (lane == 0) { rec[0] = __int_as_float(h); rec[1] = m_run; rec[2] = l_run; }
        ++sg;
    }
    if (sg == 0) { ATR(1); ATR(2); ATR(3); ATR(4); }
    ATR(5);
    cbar();
    ATR(6);
    // ---- merge the warp segments of this CTA per head, one global partial per (head, CTA)
    // (a CTA whose warps of head hh have no tokens still writes its record: max = -inf, sum = 0, which stage_x_attn weighs with 0)
    {
        const int h_first = (c * MG_CONSUMERS) / sp.wph, h_last = min(H - 1, (c * MG_CONSUMERS + MG_CONSUMERS - 1) / sp.wph);
        for (int hh = h_first + warp; hh <= h_last; hh += MG_CONSUMERS) {
            float M = -INFINITY;
            for (int s = 0; s < 2 * MG_CONSUMERS; ++s)
                if (__float_as_int(segs[s * D4]) == hh) M = fmaxf(M, segs[s * D4 + 1]);
            float L = 0.f;
            float num[(D + 31) / 32];
#pragma unroll
            for (int i = 0; i < (D + 31) / 32; ++i) num[i] = 0.f;
            for (int s = 0; s < 2 * MG_CONSUMERS; ++s) {
                if (__float_as_int(segs[s * D4]) != hh) continue;
                const float w = __expf(segs[s * D4 + 1] - M);
                L += w * segs[s * D4 + 2];
#pragma unroll
                for (int i = 0; i < (D + 31) / 32; ++i) {
                    const int d = lane + i * 32;
                    if (d < D) num[i] += w * segs[s * D4 + 4 + d];
                }
            }
            const int c0 = att_c0(hh, sp.wph);
            float* dst = P.att_ws + ((size_t)hh * P.att_maxp + (c - c0)) * D4;
#pragma unroll
            for (int i = 0; i < (D + 31) / 32; ++i) {
                const int d = lane + i * 32;
                if (d < D) dst[4 + d] = num[i];
            }
            if (lane == 0) { dst[0] = M; dst[1] = L; }
        }
    }
    ATR(7);
#undef ATR
}

// TRACE: the variant launched when MegaPlan::trace is set (GVL_MEGA_TRACE); the production variant carries no tracing code
template <int D, bool TRACE>
__global__ void __launch_bounds__(MG_THREADS, 1)
decode_mega_kernel(const MegaPlan* __restrict__ plan_g, int n_steps, long long* tokens_out, float* logits_out,
                   long long eos_id, long long pad_id) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t s_full[MG_CONSUMERS * MG_SLOTS], s_empty[MG_CONSUMERS * MG_SLOTS];
    __shared__ float s_red[MG_CONSUMERS];
    __shared__ float s_rope[2 * 128];
    __shared__ unsigned s_epoch;
    const MegaPlan& P = *plan_g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    Smem S;
    S.ring = smem;
    S.xa = smem + MG_RING_BYTES;
    S.part = reinterpret_cast<float*>(S.xa + P.x_bytes);
    S.full = s_full;
    S.empty = s_empty;
    S.red = s_red;
    S.rope = s_rope;

    if (tid == 0) {
        s_epoch = 0;
        for (int s = 0; s < MG_CONSUMERS * MG_SLOTS; ++s) {
            ptx::mbar_init(ptx::smem_u32(&s_full[s]), 1);
            ptx::mbar_init(ptx::smem_u32(&s_empty[s]), 1);
        }
        ptx::fence_mbar_init();
    }
    __syncthreads();
    const int n_ops = P.n_layers * 4 + 1;

    if (warp == MG_CONSUMERS) {
        // ------------------------------------------------------------ producer: every step's weights, in order
        uint32_t pc = 0;
        const int ppos0 = P.st->ctx_len;
        const unsigned bps = 5u * P.n_layers + 2u;     // grid barriers per step (consumer side)
        for (int stp = 0; stp < n_steps; ++stp)
            for (int i = 0; i < n_ops; ++i) {
                const MegaOp op = P.ops[i];
                // consumer epoch (grid barriers passed) while they work on this phase: 5 phases per layer, qkv first
                const int l = i >> 2, j = i & 3;
                const unsigned my_epoch = (unsigned)stp * bps + 5u * l + (j == 0 ? 0u : j + 1u);
                produce_phase(op, S, pc, MG_INFLIGHT_AHEAD, MG_INFLIGHT_CUR, &s_epoch, my_epoch, lane);
                if (j == 0 && i + 1 < n_ops) {                              // after qkv: this layer's K / V stream
                    const unsigned need = stp > 0 ? (unsigned)(stp - 1) * bps + 5u * l + 2u : 0u;
                    produce_attention<D>(P, l, ppos0 + stp, S, pc, MG_INFLIGHT_AHEAD, MG_INFLIGHT_CUR, my_epoch + 1u, lane, &s_epoch, need);
                }
            }
        return;
    }
    // ---------------------------------------------------------------- consumers
    unsigned epoch = 0;
    uint32_t cnt = 0;
    const int pos0 = P.st->ctx_len;                     // position == cache slot of the first token processed
    const int step0 = P.st->step;
    constexpr bool tracing = TRACE;
    if (tracing && tid == 0) P.trace[(size_t)blockIdx.x * MEGA_TRACE_STRIDE + MEGA_TRACE_ATT_OFF - 2] = (long long)globaltimer_ns();
    for (int stp = 0; stp < n_steps; ++stp) {
        const int pos = pos0 + stp, step = step0 + stp;
        Tracer tr{tracing ? P.trace + (size_t)blockIdx.x * MEGA_TRACE_STRIDE : nullptr};
        tr.mark(tid); tr.mark(tid); tr.mark(tid);       // (the embed phase of the first version: kept for the trace layout)
        long long tok = __ldcg(&P.st->cur_token);
        tok = tok < 0 ? 0 : (tok >= P.vocab ? P.vocab - 1 : tok);      // ids set through the C ABI are not trusted
        const __nv_bfloat16* emb_row = P.embed + (size_t)tok * P.dim;
        if (tid < P.head_dim) {
            s_rope[tid] = __bfloat162float(P.rope_cos[(size_t)pos * P.head_dim + tid]);
            s_rope[128 + tid] = __bfloat162float(P.rope_sin[(size_t)pos * P.head_dim + tid]);
        }
        MegaOp op = P.ops[0];
        NormPre np;
        prefetch_norm(op, np, tid);
        auto boundary = [&]() { grid_barrier(P.grid_bar, epoch, tid, &s_epoch); };
        // One loop over the GEMV phases (4 per layer: qkv | o_proj | gate_up | down, then lm_head), with the attention phase after
        // every qkv: stage_x_* / gemv_items / attention_phase exist ONCE in the kernel. The first version spelled the layer out
        // (5 inlined copies of the GEMV phase): 13.1k SASS instructions = 210 KB, and the ncu capture had 8 % of the warp samples on
        // stall_no_inst in the once-per-layer boundary code (profiles/r2_decode.md).
        for (int i = 0; i < n_ops; ++i) {
            const int l = i >> 2, j = i & 3;
            long long* occ = (tracing && l == P.n_layers - 1) ? P.trace + (size_t)blockIdx.x * MEGA_TRACE_STRIDE + MEGA_TRACE_OCC_OFF : nullptr;
            // input vector: merge of the attention partials (o_proj) or a vector in global memory (+ RMSNorm)
            if (op.x_kind == 1) stage_x_attn<D>(P, S, pos + 1, tid, occ ? occ + 96 + 8 : nullptr);
            else stage_x_vec(op, (op.from_embed & 1) ? emb_row : op.x, np, S, tid, warp, lane, occ ? occ + 96 + 8 * j : nullptr);
            tr.mark(tid);
            gemv_items(op, P, emb_row, (i == n_ops - 1 && logits_out) ? logits_out + (size_t)step * P.vocab : nullptr, S, cnt, tid, warp, lane,
                       occ ? occ + 8 * j : nullptr);
            if (i + 1 < n_ops) {
                op = P.ops[i + 1];
                prefetch_norm(op, np, tid);
            }
            tr.mark(tid); boundary(); tr.mark(tid);
            if (j == 0 && i + 1 < n_ops) {
                // rope + KV append + split-KV attention
                attention_phase<D, TRACE>(P, l, pos, S, cnt, warp, lane);
                tr.mark(tid);                           // keeps 3 marks per phase (no staging step here)
                tr.mark(tid); boundary(); tr.mark(tid);
            }
        }
        // ---- bookkeeping (CTA 0): tokens_out[step] = argmax (or pad after EOS), ctx_len++, step++
        if (blockIdx.x == 0 && tid == 0) {
            DecodeState* st = P.st;
            const unsigned long long key = __ldcg(P.amax);
            *P.amax = 0ull;
            long long nt = key != 0ull ? (long long)(0xffffffffu - (uint32_t)(key & 0xffffffffull)) : 0;
            if (st->finished) nt = pad_id;
            else if (eos_id >= 0 && nt == eos_id) st->finished = 1;
            if (tokens_out) tokens_out[step] = nt;
            st->cur_token = nt;
            st->ctx_len = pos + 1;
            st->attn_len = pos + 1;
            st->step = step + 1;
        }
        if (stp + 1 < n_steps) boundary();              // the next step reads cur_token
    }
    if (tracing && tid == 0) P.trace[(size_t)blockIdx.x * MEGA_TRACE_STRIDE + MEGA_TRACE_ATT_OFF - 1] = (long long)globaltimer_ns();
}

}  // namespace

bool decode_mega_shape(MegaOp* op) {
    const int K = op->K;
    if (K <= 0 || K % 64 != 0) return false;
    int nseg = (K + MEGA_SEG - 1) / MEGA_SEG;
    while (nseg <= K / 64 && (K % nseg != 0 || (K / nseg) % 64 != 0)) ++nseg;
    if (nseg > K / 64) return false;
    op->nseg = nseg;
    op->seg_len = K / nseg;
    op->n_out = op->act == 3 ? op->n_rows / 2 : op->n_rows;
    if (op->act == 3 && op->n_rows % 256 != 0) return false;
    op->units = (op->n_out + MEGA_ROWS - 1) / MEGA_ROWS;
    return true;
}


namespace {
// out: [unit][sel][seg][32-k chunk][row g][32 k]; one thread per 16-byte chunk (8 k of one row)
__global__ void mega_pack_kernel(const __nv_bfloat16* __restrict__ W, int ldw, MegaOp op, __nv_bfloat16* __restrict__ out,
                                 size_t n_vec) {
    const int nsel = op.act == 3 ? 2 : 1;
    const int cpi = op.seg_len / 32;                   // chunks per item
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < n_vec; v += (size_t)gridDim.x * blockDim.x) {
        const int t = (int)(v & 3), g = (int)((v >> 2) & 7);
        size_t rest = v >> 5;
        const int ch = (int)(rest % cpi); rest /= cpi;
        const int seg = (int)(rest % op.nseg); rest /= op.nseg;
        const int sel = (int)(rest % nsel);
        const int unit = (int)(rest / nsel);
        const int n = row_base_of(op, unit, sel) + g;
        uint4 val = make_uint4(0u, 0u, 0u, 0u);
        if (n < op.n_rows) val = *reinterpret_cast<const uint4*>(W + (size_t)n * ldw + (size_t)seg * op.seg_len + ch * 32 + t * 8);
        reinterpret_cast<uint4*>(out)[v] = val;
    }
}
}  // namespace

size_t decode_mega_packed_elems(const MegaOp* op) {
    return (size_t)op->units * (op->act == 3 ? 2 : 1) * op->nseg * MEGA_ROWS * op->seg_len;
}

int decode_mega_pack(const MegaOp* op, const __nv_bfloat16* W, int ldw, __nv_bfloat16* dst, cudaStream_t s) {
    if (ldw % 8 != 0 || (reinterpret_cast<uintptr_t>(W) & 15) != 0) return GVL_ERR_ALIGN;
    const size_t n_vec = decode_mega_packed_elems(op) / 8;
    mega_pack_kernel<<<num_sms() * 8, 256, 0, s>>>(W, ldw, *op, dst, n_vec);
    return cudaGetLastError() == cudaSuccess ? GVL_OK : GVL_ERR_CUDA;
}

bool decode_mega_finalize(MegaPlan* p) {
    const int G = num_sms();
    const int D = p->head_dim, H = p->heads;
    if (D != 64 && D != 96 && D != 128) return false;
    if (H > 64 || H % p->kv_heads != 0 || (H * D) % 64 != 0) return false;
    int maxk = 0, items = 0;
    const int n_ops = p->n_layers * 4 + 1;
    for (int i = 0; i < n_ops; ++i) {
        const MegaOp& op = p->ops[i];
        maxk = op.K > maxk ? op.K : maxk;
        const int nu = (op.units + G - 1) / G;
        const int it = nu * (op.act == 3 ? 2 : 1) * op.nseg;
        items = it > items ? it : items;
    }
    int xb = maxk * 2;
    if (att_scratch_bytes(D) > xb) xb = att_scratch_bytes(D);
    p->x_bytes = (xb + 127) & ~127;
    {
        // partial-sum buffer: the largest phase in one pass if it fits; otherwise what is left of shared memory (gemv_items then runs
        // that phase in passes). At least 64 items (8 units of 8 segments) must fit.
        const long avail = (long)MG_SMEM_LIMIT - MG_RING_BYTES - p->x_bytes;
        long cap = avail / 32;
        if (cap < 64) return false;
        p->part_items = items < cap ? items : (int)cap;
        if (p->part_items < 64) p->part_items = 64;          // stage_x_attn keeps its 2 KB of softmax weights here
    }
    p->att_maxp = G / H + 2;
    return (size_t)MG_RING_BYTES + p->x_bytes + (size_t)p->part_items * 32 <= (size_t)MG_SMEM_LIMIT;
}

size_t decode_mega_att_ws_bytes(const MegaPlan* p) {
    return (size_t)p->heads * p->att_maxp * (p->head_dim + 4) * sizeof(float);
}

void decode_mega_attention_split(int ctx, int heads, int n_ctas, int* warps_per_head, int* tokens_per_warp, int* max_partials) {
    const AttSplit a = att_split(ctx, heads, n_ctas);
    *warps_per_head = a.wph;
    *tokens_per_warp = a.lw;
    int mp = 0;
    for (int h = 0; h < heads; ++h) mp = max(mp, att_c1(h, a.wph) - att_c0(h, a.wph) + 1);
    *max_partials = mp;
}

int decode_mega_launch(const MegaPlan* hp, const MegaPlan* plan_dev, int n_steps, long long* tokens_out, float* logits_out,
                       long long eos_id, long long pad_id, cudaStream_t s) {
    const size_t smem = (size_t)MG_RING_BYTES + hp->x_bytes + (size_t)hp->part_items * 32;
    using KernelT = void (*)(const MegaPlan*, int, long long*, float*, long long, long long);
    const bool tr = hp->trace != nullptr;
    const int di = (hp->head_dim == 64 ? 0 : hp->head_dim == 96 ? 1 : 2) + (tr ? 3 : 0);
    static const KernelT kerns[6] = {decode_mega_kernel<64, false>, decode_mega_kernel<96, false>, decode_mega_kernel<128, false>,
                                     decode_mega_kernel<64, true>,  decode_mega_kernel<96, true>,  decode_mega_kernel<128, true>};
    KernelT kern = kerns[di];
    static size_t attr_set_tab[64][6] = {};            // per device (cudaFuncSetAttribute is per device), per kernel variant
    int dev = 0;
    cudaGetDevice(&dev);
    size_t* attr_set = attr_set_tab[dev & 63];
    if (attr_set[di] < smem) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return GVL_ERR_CUDA;
        attr_set[di] = smem;
    }
    if (cudaMemsetAsync(hp->grid_bar, 0, sizeof(unsigned), s) != cudaSuccess) return GVL_ERR_CUDA;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(num_sms());
    cfg.blockDim = dim3(MG_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, kern, plan_dev, n_steps, tokens_out, logits_out, eos_id, pad_id) != cudaSuccess)
        return GVL_ERR_CUDA;
    g_launch_count++;
    return cudaGetLastError() == cudaSuccess ? GVL_OK : GVL_ERR_CUDA;
}

}  // namespace gvl
