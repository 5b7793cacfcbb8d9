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
// padding. The warp multiplies the item with mma.sync.m16n8k16 (A = the activation vector broadcast over the
// 16 rows, B = the 8 weight rows; k is permuted identically on both operands so every lane feeds its fragments with
// 128-bit shared-memory loads). That is ~6 warp instructions per KB of weights, so the consumers drain the ring several
// times faster than HBM fills it: the ring runs near-empty inside a phase and absorbs ~4 us of HBM stream while the
// consumers sit in a barrier / stage the next activation vector. Items of a CTA are dealt round-robin to its 8 warps
// (k-split inside the CTA, partial sums reduced through shared memory in a fixed order -> deterministic).
// Every consumer warp has its own 3-slot ring fed by one lane of the producer warp, which keeps at most `inflight`
// copies per lane outstanding: 8 x 8 KB x 148 SMs = 9.5 MB in flight is enough for the full HBM rate, while deeper
// queues (the first version had 28 MB outstanding) only add queueing delay (4+ us measured) to every latency-critical
// load of the phase boundaries (activation staging, barrier atomics, q/k of the attention phase).
//
// Attention phase: the flat (head, token) space is cut into equal contiguous ranges, one per consumer warp of the
// grid; a warp streams K/V rows (4 lanes per token, 8 tokens per pass, 4 passes of K and V loads in flight) with an
// online softmax, warps of a CTA merge through shared memory, and the <= G/H + 2 CTA partials per head are merged by
// every CTA while it stages the o_proj input (no atomics, no extra barrier).
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
constexpr int MG_THREADS = 32 * (MG_CONSUMERS + 2);          // 8 consumer warps, ring producer, L2 prefetcher
constexpr int MG_SLOT_BYTES = MEGA_ROWS * MEGA_SEG * 2;          // 8192: one packed item
constexpr int MG_SLOTS = 3;                                      // ring depth per consumer warp
constexpr int MG_RING_BYTES = MG_CONSUMERS * MG_SLOTS * MG_SLOT_BYTES;   // 196608
constexpr int MG_SMEM_LIMIT = 232448 - 2048;                     // 227 KB opt-in minus static shared memory
constexpr int MG_ATT_SHORT = 128;                                // ctx <= this: one warp per head

__host__ __device__ inline int att_warp_len(int ctx, int H, int G) {
    if (ctx <= MG_ATT_SHORT) return ctx;
    const int TW = G * MG_CONSUMERS;
    const int per = (H * ctx + TW - 1) / TW;
    return (per + 7) & ~7;
}
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
__device__ __forceinline__ void l2_prefetch(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<uint64_t>(src)), "r"(bytes) : "memory");
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
__device__ __forceinline__ uint4 ldg_stream4(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
    return r;
}
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cbar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }   // the 256 consumer threads

// consumers-only grid barrier
// `epoch` counts phase boundaries (published for the producer warp), `bars` counts the REAL grid barriers among them (in the
// flags-in-data mode most boundaries are phase_end() below)
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& epoch, unsigned& bars, int tid,
                                             volatile unsigned* cta_epoch, int ablate = 0) {
    cbar();
    ++epoch;
    if (ablate & 1) {                                // timing ablation only (results are wrong)
        if (tid == 0) *cta_epoch = epoch;
        return;
    }
    ++bars;
    if (tid == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        const unsigned target = bars * gridDim.x;
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
// flags-in-data mode: a phase boundary without a grid barrier (the next phase polls its inputs)
__device__ __forceinline__ void phase_end(unsigned& epoch, int tid, volatile unsigned* cta_epoch) {
    cbar();
    ++epoch;
    if (tid == 0) {
        __threadfence_block();
        *cta_epoch = epoch;
    }
}

// ------------------------------------------------------------------ flags-in-data packets
__device__ __forceinline__ uint4 ld_poll4(const void* p) {          // L1-bypassing, never hoisted out of a polling loop
    uint4 r;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ uint2 ld_poll2(const void* p) {
    uint2 r;
    asm volatile("ld.global.cg.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void st_packet2(uint2* p, uint32_t data, uint32_t flag) {
    asm volatile("st.global.cg.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(data), "r"(flag) : "memory");
}
__device__ __forceinline__ void st_packet4(uint4* p, float a, float b, float c, uint32_t flag) {
    asm volatile("st.global.cg.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(__float_as_uint(a)), "r"(__float_as_uint(b)),
                 "r"(__float_as_uint(c)), "r"(flag) : "memory");
}
struct PollGuard {                                                  // a producer died or the phase ids are out of step: trap, never hang
    long long t0;
    __device__ __forceinline__ PollGuard() : t0(0) {}
    __device__ __forceinline__ void spin(const char* what, unsigned want, unsigned got) {
        if (t0 == 0) { t0 = clock64(); __nanosleep(200); return; }
        __nanosleep(400);                                           // 38k threads poll: keep the request rate off the weight stream
        if (clock64() - t0 > 4000000000LL) {
            printf("gvl: decode_mega %s poll timeout block %d thread %d want %u got %u\n", what, blockIdx.x, threadIdx.x, want, got);
            __trap();
        }
    }
};
// 4 consecutive elements (two packets) of an LL vector, polled until both carry `flag`; returns the 4 bf16 as (lo pair, hi pair)
__device__ __forceinline__ uint2 poll_vec4(const uint2* v, int idx4, unsigned flag, const char* what) {
    PollGuard g;
    uint4 r = ld_poll4(v + (size_t)idx4 * 2);
    while (r.y != flag || r.w != flag) {
        g.spin(what, flag, r.y != flag ? r.y : r.w);
        r = ld_poll4(v + (size_t)idx4 * 2);
    }
    return make_uint2(r.x, r.z);
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
                                            int tid, int warp, int lane) {
    __nv_bfloat16* sx = reinterpret_cast<__nv_bfloat16*>(S.xa);
    const int kv = op.K / 8;
    if (op.norm_w == nullptr) {
        for (int i = tid; i < kv; i += 256) reinterpret_cast<uint4*>(sx)[i] = ldcg4(xsrc + (size_t)i * 8);
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
        cbar();
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < MG_CONSUMERS; ++w) t += S.red[w];
        const float rstd = rsqrtf(t / op.K + op.eps);
#pragma unroll
        for (int u = 0; u < 2; ++u)
            if (tid + u * 256 < kv) reinterpret_cast<uint4*>(sx)[tid + u * 256] = norm8(xv[u], np.v[u], rstd);
        cbar();
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
        reinterpret_cast<uint4*>(sx)[i] = norm8(reinterpret_cast<uint4*>(sx)[i], __ldg(reinterpret_cast<const uint4*>(op.norm_w) + i), rstd);
    cbar();
}

// ------------------------------------------------------------------ x staging from {2 x bf16, flag} packets (+ RMSNorm): the poll IS the load
__device__ __noinline__ void stage_x_ll(const MegaOp& op, unsigned flag, const Smem& S, int tid, int warp, int lane) {
    __nv_bfloat16* sx = reinterpret_cast<__nv_bfloat16*>(S.xa);
    const int nq = op.K / 4;                                       // 4 elements per thread-load; K <= 8192 -> <= 8 per thread
    uint2 nw[8];
    if (op.norm_w != nullptr) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (tid + u * 256 < nq) nw[u] = __ldg(reinterpret_cast<const uint2*>(op.norm_w) + tid + u * 256);
    }
    // 38k threads re-polling every packet they need put more requests on the L2 than the weight stream itself: only warp 0
    // waits, on 32 sentinel packets spread over the vector (32 different producer CTAs), everybody else sleeps at the CTA barrier;
    // afterwards every thread loads its packets once, still verifying the flags (a late packet is re-polled individually)
    if (warp == 0) {
        PollGuard g;
        const uint2* sp = op.x_ll + (size_t)(lane * (nq / 32)) * 2;
        uint4 r = ld_poll4(sp);
        while (!__all_sync(0xffffffffu, r.y == flag && r.w == flag)) {
            g.spin("x sentinel", flag, r.y);
            r = ld_poll4(sp);
        }
    }
    cbar();
    // all of this thread's packet loads are issued before the first flag is examined (one L2 round trip for the whole vector when
    // the data is there; a per-packet poll loop would serialise up to 8 round trips); late packets are re-polled individually
    uint4 raw[8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
        if (tid + u * 256 < nq) raw[u] = ld_poll4(op.x_ll + (size_t)(tid + u * 256) * 2);
    uint2 xv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
        if (tid + u * 256 < nq) {
            PollGuard g;
            while (raw[u].y != flag || raw[u].w != flag) {
                g.spin("x", flag, raw[u].y != flag ? raw[u].y : raw[u].w);
                raw[u] = ld_poll4(op.x_ll + (size_t)(tid + u * 256) * 2);
            }
            xv[u] = make_uint2(raw[u].x, raw[u].z);
        }
    if (op.norm_w == nullptr) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (tid + u * 256 < nq) reinterpret_cast<uint2*>(sx)[tid + u * 256] = xv[u];
        cbar();
        return;
    }
    float ss = 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u)
        if (tid + u * 256 < nq) {
            float2 f;
            f = unpack_bf16(xv[u].x); ss += f.x * f.x + f.y * f.y;
            f = unpack_bf16(xv[u].y); ss += f.x * f.x + f.y * f.y;
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) S.red[warp] = ss;
    cbar();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < MG_CONSUMERS; ++w) t += S.red[w];
    const float rstd = rsqrtf(t / op.K + op.eps);
#pragma unroll
    for (int u = 0; u < 8; ++u)
        if (tid + u * 256 < nq) {
            float2 f, g;
            uint2 o;
            f = unpack_bf16(xv[u].x); g = unpack_bf16(nw[u].x); o.x = pack_bf16(bf16r(f.x * rstd) * g.x, bf16r(f.y * rstd) * g.y);
            f = unpack_bf16(xv[u].y); g = unpack_bf16(nw[u].y); o.y = pack_bf16(bf16r(f.x * rstd) * g.x, bf16r(f.y * rstd) * g.y);
            reinterpret_cast<uint2*>(sx)[tid + u * 256] = o;
        }
    cbar();
}

// merge of the split-KV partials from {3 floats, flag} packets: task (head, l) owns dims l, l + 32, l + 64 of that head
__device__ __noinline__ void stage_x_attn_ll(const MegaPlan& P, unsigned flag, const Smem& S, int ctx, int tid) {
    __nv_bfloat16* sx = reinterpret_cast<__nv_bfloat16*>(S.xa);
    const int D = P.head_dim, H = P.heads, maxp = P.att_maxp;
    const int Lc = att_warp_len(ctx, H, gridDim.x) * MG_CONSUMERS;
    if (tid < 32) {                                                 // sentinels: the (m, l) packet of every partial, one head per lane
        for (int h = tid; h < H; h += 32) {
            const int c0 = (h * ctx) / Lc, c1 = ((h + 1) * ctx - 1) / Lc;
            for (int s = 0; s <= c1 - c0; ++s) {
                PollGuard g;
                const uint4* sp = P.att_ll + ((size_t)h * maxp + s) * 33 + 32;
                uint4 r = ld_poll4(sp);
                while (r.w != flag) { g.spin("att sentinel", flag, r.w); r = ld_poll4(sp); }
            }
        }
    }
    cbar();
    for (int task = tid; task < H * 32; task += 256) {
        const int h = task >> 5, l = task & 31;
        const int c0 = (h * ctx) / Lc, c1 = ((h + 1) * ctx - 1) / Lc;
        const int np = c1 - c0 + 1;
        const uint4* base = P.att_ll + (size_t)h * maxp * 33;
        // all partials of the head in flight at once (np <= 8 in this mode), re-polled individually until complete
        uint4 hd[8], ov[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (j < np) { hd[j] = ld_poll4(base + (size_t)j * 33 + 32); ov[j] = ld_poll4(base + (size_t)j * 33 + l); }
        float M = -INFINITY;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (j < np) {
                PollGuard g;
                while (hd[j].w != flag || ov[j].w != flag) {
                    g.spin("att", flag, hd[j].w != flag ? hd[j].w : ov[j].w);
                    hd[j] = ld_poll4(base + (size_t)j * 33 + 32);
                    ov[j] = ld_poll4(base + (size_t)j * 33 + l);
                }
                M = fmaxf(M, __uint_as_float(hd[j].x));
            }
        float den = 0.f, n0 = 0.f, n1 = 0.f, n2 = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (j < np) {
                const float w = __expf(__uint_as_float(hd[j].x) - M);
                den += w * __uint_as_float(hd[j].y);
                n0 += w * __uint_as_float(ov[j].x); n1 += w * __uint_as_float(ov[j].y); n2 += w * __uint_as_float(ov[j].z);
            }
        const float inv = den > 0.f ? 1.0f / den : 0.f;
        sx[h * D + l] = __float2bfloat16_rn(n0 * inv);
        if (l + 32 < D) sx[h * D + l + 32] = __float2bfloat16_rn(n1 * inv);
        if (l + 64 < D) sx[h * D + l + 64] = __float2bfloat16_rn(n2 * inv);
    }
    cbar();
}

// ------------------------------------------------------------------ x staging: merge of the split-KV attention partials
__device__ __forceinline__ void stage_x_attn(const MegaPlan& P, const Smem& S, int ctx, int tid) {
    __nv_bfloat16* sx = reinterpret_cast<__nv_bfloat16*>(S.xa);
    const int D = P.head_dim, H = P.heads, D4 = D + 4, maxp = P.att_maxp;
    const int Lc = att_warp_len(ctx, H, gridDim.x) * MG_CONSUMERS;
    for (int gi = tid; gi < H * D / 4; gi += 256) {
        const int e0 = gi * 4, h = e0 / D, d0 = e0 - h * D;
        const int c0 = (h * ctx) / Lc, c1 = ((h + 1) * ctx - 1) / Lc;
        const int np = c1 - c0 + 1;
        const float* base = P.att_ws + (size_t)h * maxp * D4;
        float4 num = make_float4(0.f, 0.f, 0.f, 0.f);
        float den = 0.f;
        if (np <= 8) {
            float2 hd[8];
            float4 ov[8];
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                if (s < np) {
                    hd[s] = __ldcg(reinterpret_cast<const float2*>(base + (size_t)s * D4));
                    ov[s] = __ldcg(reinterpret_cast<const float4*>(base + (size_t)s * D4 + 4 + d0));
                } else {
                    hd[s] = make_float2(-INFINITY, 0.f);
                    ov[s] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            float M = hd[0].x;
#pragma unroll
            for (int s = 1; s < 8; ++s) M = fmaxf(M, hd[s].x);
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                const float w = __expf(hd[s].x - M);
                den += w * hd[s].y;
                num.x += w * ov[s].x; num.y += w * ov[s].y; num.z += w * ov[s].z; num.w += w * ov[s].w;
            }
        } else {
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
        uint2 o;
        o.x = pack_bf16(num.x * inv, num.y * inv);
        o.y = pack_bf16(num.z * inv, num.w * inv);
        *reinterpret_cast<uint2*>(sx + e0) = o;
    }
    cbar();
}

// ------------------------------------------------------------------ consumer side of one GEMV phase (x already staged)
template <bool LL>
__device__ __forceinline__ void gemv_items(const MegaOp& op, const MegaPlan& P, const __nv_bfloat16* emb_row,
                                           float* extra_out, const Smem& S, uint32_t& cnt, int tid, int warp, int lane,
                                           long long* occ = nullptr, unsigned out_flag = 0u) {
    const int G = gridDim.x, c = blockIdx.x;
    const int nu = op.units > c ? (op.units - c + G - 1) / G : 0;
    const int nsel = op.act == 3 ? 2 : 1;
    const int nseg = op.nseg;
    const int ipu = nsel * nseg;
    const int n_items = nu * ipu;
    const int g = lane >> 2, t = lane & 3;
    const int nchunk2 = op.seg_len / 64;
    // residual of this thread's output column: loaded now, used in the epilogue (hides one L2 round trip)
    constexpr bool ll = LL;
    const bool res_is_ll = ll && !(op.from_embed & 2) && op.res_ll != nullptr;
    const __nv_bfloat16* res = (op.from_embed & 2) ? emb_row : (res_is_ll ? reinterpret_cast<const __nv_bfloat16*>(op.res_ll) : op.residual);
    // element n of a residual vector: plain bf16, or the data half of its {2 x bf16, flag} packet (this CTA polled the whole vector
    // when it staged the phase that consumed it, so no flag check here)
    auto load_res = [&](int n) -> float {
        if (res_is_ll) {
            const uint32_t d = __ldcg(reinterpret_cast<const unsigned int*>(op.res_ll + (n >> 1)));
            return __uint_as_float((n & 1) ? (d & 0xffff0000u) : (d << 16));
        }
        return __uint_as_float((uint32_t)__ldcg(reinterpret_cast<const unsigned short*>(res) + n) << 16);
    };
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
    const uint32_t ring_w = ptx::smem_u32(S.ring) + warp * (MG_SLOTS * MG_SLOT_BYTES) + g * 64 + t * 16;   // item: [chunk][row g][64 B]
    const uint32_t x_u32 = ptx::smem_u32(S.xa) + t * 16;
    for (int q = warp; q < n_items; q += MG_CONSUMERS) {
        const uint32_t slot = cnt % MG_SLOTS, par = (cnt / MG_SLOTS) & 1;
        ptx::mbar_wait(ptx::smem_u32(&S.full[warp * MG_SLOTS + slot]), par);
        const int j = q / ipu, r = q - j * ipu;
        const int seg = r % nseg;
        uint32_t wa = ring_w + slot * MG_SLOT_BYTES;
        uint32_t xa = x_u32 + seg * op.seg_len * 2;
        float acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a) acc[a][0] = acc[a][1] = acc[a][2] = acc[a][3] = 0.f;
#pragma unroll 2
        for (int ch = 0; ch < ((P.ablate & 8) ? 0 : nchunk2); ++ch) {
            const uint4 w0 = lds128(wa), x0 = lds128(xa);
            const uint4 w1 = lds128(wa + 512), x1 = lds128(xa + 64);
            mma16816(acc[0], x0.x, x0.x, x0.y, x0.y, w0.x, w0.y);
            mma16816(acc[1], x0.z, x0.z, x0.w, x0.w, w0.z, w0.w);
            mma16816(acc[2], x1.x, x1.x, x1.y, x1.y, w1.x, w1.y);
            mma16816(acc[3], x1.z, x1.z, x1.w, x1.w, w1.z, w1.w);
            wa += 1024;
            xa += 128;
        }
        if (g == 0) {
            float2 o;
            o.x = (acc[0][0] + acc[1][0]) + (acc[2][0] + acc[3][0]);
            o.y = (acc[0][1] + acc[1][1]) + (acc[2][1] + acc[3][1]);
            *reinterpret_cast<float2*>(S.part + (size_t)q * 8 + 2 * t) = o;
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&S.empty[warp * MG_SLOTS + slot]));
        ++cnt;
    }
    cbar();
    // ---- epilogue: one thread per output column, segments summed in a fixed order
    unsigned long long key = 0ull;
    const bool out_ll = ll && op.out_ll != nullptr;
    for (int o0 = 0; o0 < nu * MEGA_ROWS; o0 += 256) {           // warp-uniform trip count: the packet store pairs lanes
        const int o = o0 + tid;
        const bool active = o < nu * MEGA_ROWS;
        const int j = o >> 3, col = o & 7;
        const int n = (c + j * G) * MEGA_ROWS + col;
        const bool live = active && n < op.n_out;
        unsigned short ybits = 0;
        if (live) {
            const float* pp = S.part + (size_t)j * ipu * 8 + col;
            float a0 = 0.f;
            for (int s = 0; s < nseg; ++s) a0 += pp[s * 8];
            if (op.act == 3) {
                float a1 = 0.f;
                for (int s = 0; s < nseg; ++s) a1 += pp[(nseg + s) * 8];
                const float gt = bf16r(a0), u = bf16r(a1);
                const __nv_bfloat16 yv = __float2bfloat16_rn(u * bf16r(silu_f(gt)));
                ybits = __bfloat16_as_ushort(yv);
                if (!out_ll) reinterpret_cast<__nv_bfloat16*>(op.out)[n] = yv;
            } else {
                float y = a0;
                if (op.bias) y += __bfloat162float(op.bias[n]);
                y = bf16r(y);
                if (res) {
                    const float rv = o < 256 ? res_pre : load_res(n);
                    y = bf16r(y + rv);
                }
                ybits = __bfloat16_as_ushort(__float2bfloat16_rn(y));
                if (!out_ll) {
                    if (op.out_f32) reinterpret_cast<float*>(op.out)[n] = y;
                    else reinterpret_cast<__nv_bfloat16*>(op.out)[n] = __float2bfloat16_rn(y);
                }
                if (extra_out) extra_out[n] = y;
                if (op.argmax) {
                    uint32_t u = __float_as_uint(y);
                    u ^= (u >> 31) ? 0xffffffffu : 0x80000000u;
                    const unsigned long long k = ((unsigned long long)u << 32) | (unsigned long long)(0xffffffffu - (uint32_t)n);
                    key = k > key ? k : key;
                }
            }
        }
        if (out_ll) {
            // columns n, n + 1 (n even) of one 8-row unit sit in adjacent lanes: one 8-byte {2 x bf16, phase id} packet
            const unsigned hi = __shfl_down_sync(0xffffffffu, (unsigned)ybits, 1);
            if (live && !(col & 1)) st_packet2(op.out_ll + (n >> 1), (unsigned)ybits | (hi << 16), out_flag);
        }
    }
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
__device__ __forceinline__ int cta_items(const MegaOp& op) {
    const int G = gridDim.x, c = blockIdx.x;
    const int nu = op.units > c ? (op.units - c + G - 1) / G : 0;
    return nu * (op.act == 3 ? 2 : 1) * op.nseg;
}

// `inflight` copies per lane while the phase is still ahead of the consumers (pure prefetch: deep queues only delay the
// latency-critical loads of the phase boundary the consumers are in), `inflight_cur` once the consumers have reached
// this phase (cta_epoch >= my_epoch) and are waiting for exactly these items.
__device__ __forceinline__ void produce_phase(const MegaOp& op, const Smem& S, uint32_t& pc, int inflight, int inflight_cur,
                                              volatile unsigned* cta_epoch, unsigned my_epoch, int lane,
                                              volatile unsigned* prod_ord, unsigned ord_base) {
    const int G = gridDim.x, c = blockIdx.x;
    const int nu = op.units > c ? (op.units - c + G - 1) / G : 0;
    const int ipu = (op.act == 3 ? 2 : 1) * op.nseg;
    const int n_items = nu * ipu;
    if (lane == 0) *prod_ord = ord_base;               // progress in this CTA's item order, read by the L2 prefetcher
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
                if (lane == 0) *prod_ord = ord_base + (unsigned)m * MG_CONSUMERS;
            }
        }
        if (!__any_sync(0xffffffffu, issued)) __nanosleep(40);
    }
}

// ------------------------------------------------------------------ L2 prefetcher (10th warp)
// The rings hold 28 MB, i.e. ~4 us of HBM stream, but a phase boundary (epilogue stores -> grid barrier -> activation
// staging) lasts longer than that, so with the rings alone HBM idles at every boundary. This warp walks the same item
// order as the producer and issues cp.async.bulk.prefetch.L2 for items up to `win` items (8 KB each) ahead of it: HBM ->
// L2 keeps streaming while the rings are full, and after the boundary the ring refills from L2 instead of HBM.
__device__ __forceinline__ void pf_throttle(volatile unsigned* prod_ord, unsigned my_ord, int win) {
    while ((int)(my_ord - *prod_ord) > win) __nanosleep(200);
}
__device__ __forceinline__ void prefetch_phase(const MegaOp& op, volatile unsigned* prod_ord, unsigned ord_base, int win, int lane) {
    const int G = gridDim.x, c = blockIdx.x;
    const int ipu = (op.act == 3 ? 2 : 1) * op.nseg;
    const int n_items = cta_items(op);
    const uint32_t item_bytes = (uint32_t)op.seg_len * 2 * MEGA_ROWS;
    for (int q0 = 0; q0 < n_items; q0 += 32) {
        if ((int)(ord_base + q0 + 32 - *prod_ord) <= 0) continue;      // the producer is already past this batch
        pf_throttle(prod_ord, ord_base + q0, win);
        const int q = q0 + lane;
        if (q < n_items) {
            const int j = q / ipu, r = q - j * ipu;
            const size_t item = (size_t)(c + j * G) * ipu + r;
            l2_prefetch(reinterpret_cast<const uint8_t*>(op.W) + item * item_bytes, item_bytes);
        }
    }
}
template <int D>
__device__ __forceinline__ void prefetch_attention(const MegaPlan& P, int layer, int pos, volatile unsigned* prod_ord,
                                                   unsigned ord_base, int win, int lane) {
    pf_throttle(prod_ord, ord_base, win);
    if (lane >= MG_CONSUMERS) return;
    const int H = P.heads, KVH = P.kv_heads, rep = H / KVH;
    const int ctx = pos + 1;
    const int Lw = att_warp_len(ctx, H, gridDim.x);
    const int total = H * ctx;
    const __nv_bfloat16* kc_l = P.kv + (size_t)layer * 2 * KVH * P.max_ctx * D;
    const __nv_bfloat16* vc_l = kc_l + (size_t)KVH * P.max_ctx * D;
    int f = (blockIdx.x * MG_CONSUMERS + lane) * Lw;
    const int f1 = min(f + Lw, total);
    while (f < f1) {
        const int h = f / ctx;
        const int t0 = f - h * ctx;
        const int t1 = min(ctx, t0 + (f1 - f));
        f += t1 - t0;
        const int te = min(t1, pos);
        if (te > t0) {
            const size_t off = ((size_t)(h / rep) * P.max_ctx + t0) * D;
            l2_prefetch(kc_l + off, (uint32_t)(te - t0) * D * 2);
            l2_prefetch(vc_l + off, (uint32_t)(te - t0) * D * 2);
        }
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
    const int Lw = att_warp_len(ctx, H, G);
    const int total = H * ctx;
    const __nv_bfloat16* kc_l = P.kv + (size_t)layer * 2 * KVH * P.max_ctx * D;
    const __nv_bfloat16* vc_l = kc_l + (size_t)KVH * P.max_ctx * D;
    int f = 0, f1 = 0;
    if (lane < MG_CONSUMERS) {
        f = (c * MG_CONSUMERS + lane) * Lw;
        f1 = min(f + Lw, total);
    }
    int tb = 0, tend = 0, cs = 8, hk = 0, is_v = 0;
    auto next_seg = [&]() -> bool {
        while (f < f1) {
            const int h = f / ctx;
            const int t0 = f - h * ctx;
            const int t1 = min(ctx, t0 + (f1 - f));
            f += t1 - t0;
            const int te = min(t1, pos);                 // the new token (pos) never comes from the cache
            if (te > t0) {
                tb = t0; tend = te; hk = h / rep; is_v = 0;
                cs = att_chunk_len(te - t0, CTM);
                return true;
            }
        }
        return false;
    };
    bool active = next_seg();
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
                    if (tb >= tend) active = next_seg();
                } else {
                    is_v = 1;
                }
            }
        }
        if (!__any_sync(0xffffffffu, issued)) __nanosleep(40);
    }
}

// ------------------------------------------------------------------ attention phase
template <int D, bool LL>
__device__ __forceinline__ void attention_phase(const MegaPlan& P, int layer, int pos, const Smem& S, uint32_t& cnt, int warp,
                                                int lane, unsigned in_flag = 0u, unsigned out_flag = 0u) {
    constexpr bool ll = LL;
    constexpr int EPL = D / 4, VPL = EPL / 8, half = D / 2, D4 = D + 4;
    constexpr int ROWB = D * 2, CTM = (MG_SLOT_BYTES / ROWB) & ~7, NPASS = CTM / 8;   // tokens / 8-token passes per ring item
    const int H = P.heads, KVH = P.kv_heads, rep = H / KVH;
    const int G = gridDim.x, c = blockIdx.x;
    const int ctx = pos + 1;
    const int Lw = att_warp_len(ctx, H, G), Lc = Lw * MG_CONSUMERS;
    const int total = H * ctx;
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
    long long* atr = (P.trace && layer == P.n_layers - 1 && lane == 0)
                         ? P.trace + (size_t)blockIdx.x * MEGA_TRACE_STRIDE + MEGA_TRACE_ATT_OFF + warp * 8 : nullptr;
#define ATR(k) do { if (atr) atr[k] = clock64(); } while (0)
    ATR(0);
    if (lane < 2) segs[(warp * 2 + lane) * D4] = __int_as_float(-1);
    const int f0 = (c * MG_CONSUMERS + warp) * Lw;
    const int f1 = min(f0 + Lw, total);
    int f = f0, sg = 0;
    while (f < f1) {
        const int h = f / ctx, hk = h / rep;
        const int t0 = f - h * ctx;
        const int t1 = min(ctx, t0 + (f1 - f));
        f += t1 - t0;
        // ---- RoPE of q (and of the new k) for this head: q_embed = bf16(bf16(q*cos) + bf16(rot(q)*sin))
        __syncwarp();
        {
            const __nv_bfloat16* qh = qkv + (size_t)h * D;
            const __nv_bfloat16* kh = qkv + (size_t)(H + hk) * D;
            float q1[(half + 31) / 32], q2[(half + 31) / 32], k1[(half + 31) / 32], k2[(half + 31) / 32];
            if (ll) {
                // q / k of this head as {2 x bf16, phase id} packets written by the qkv phase: all loads first, then re-poll
                // whatever is not there yet (the poll replaces the grid barrier AND the load that followed it)
                constexpr int NU = (half + 31) / 32;
                const int eq = h * D, ek = (H + hk) * D;
                if (lane < 3) {                                     // sentinels: last packet of this head's q, k and v rows
                    const int e = (lane == 0 ? eq : lane == 1 ? ek : (H + KVH + hk) * D) + D - 2;
                    PollGuard g;
                    uint2 r = ld_poll2(P.qkv_ll + (e >> 1));
                    while (r.y != in_flag) { g.spin("qkv sentinel", in_flag, r.y); r = ld_poll2(P.qkv_ll + (e >> 1)); }
                }
                __syncwarp();
                uint2 pq1[NU], pq2[NU], pk1[NU], pk2[NU];
#pragma unroll
                for (int u = 0; u < NU; ++u) {
                    const int j = lane + u * 32;
                    if (j < half) {
                        pq1[u] = ld_poll2(P.qkv_ll + ((eq + j) >> 1)); pq2[u] = ld_poll2(P.qkv_ll + ((eq + j + half) >> 1));
                        pk1[u] = ld_poll2(P.qkv_ll + ((ek + j) >> 1)); pk2[u] = ld_poll2(P.qkv_ll + ((ek + j + half) >> 1));
                    }
                }
#pragma unroll
                for (int u = 0; u < NU; ++u) {
                    const int j = lane + u * 32;
                    if (j < half) {
                        PollGuard g;
                        while (pq1[u].y != in_flag || pq2[u].y != in_flag || pk1[u].y != in_flag || pk2[u].y != in_flag) {
                            g.spin("qkv", in_flag, pq1[u].y);
                            pq1[u] = ld_poll2(P.qkv_ll + ((eq + j) >> 1)); pq2[u] = ld_poll2(P.qkv_ll + ((eq + j + half) >> 1));
                            pk1[u] = ld_poll2(P.qkv_ll + ((ek + j) >> 1)); pk2[u] = ld_poll2(P.qkv_ll + ((ek + j + half) >> 1));
                        }
                        auto pick = [](uint2 pkt, int e) { return __uint_as_float((e & 1) ? (pkt.x & 0xffff0000u) : (pkt.x << 16)); };
                        q1[u] = pick(pq1[u], eq + j); q2[u] = pick(pq2[u], eq + j + half);
                        k1[u] = pick(pk1[u], ek + j); k2[u] = pick(pk2[u], ek + j + half);
                    }
                }
            } else {
#pragma unroll
                for (int u = 0; u < (half + 31) / 32; ++u) {
                    const int j = lane + u * 32;
                    if (j < half) {
                        q1[u] = __bfloat162float(__ldcg(qh + j)); q2[u] = __bfloat162float(__ldcg(qh + j + half));
                        k1[u] = __bfloat162float(__ldcg(kh + j)); k2[u] = __bfloat162float(__ldcg(kh + j + half));
                    }
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
        // 8 consecutive elements of the new v row from packets (element index e0, multiple of 8): two 16-byte loads, four flags
        auto v8_ll = [&](int e0) -> uint4 {
            const uint2* src = P.qkv_ll + (e0 >> 1);
            PollGuard g;
            uint4 a = ld_poll4(src), b = ld_poll4(src + 2);
            while (a.y != in_flag || a.w != in_flag || b.y != in_flag || b.w != in_flag) {
                g.spin("v", in_flag, a.y);
                a = ld_poll4(src); b = ld_poll4(src + 2);
            }
            return make_uint4(a.x, a.z, b.x, b.z);
        };
        if (has_new) {
#pragma unroll
            for (int i = 0; i < VPL; ++i)
                vn[i] = ll ? v8_ll((H + KVH + hk) * D + sub * EPL + i * 8) : ldcg4(vnew + sub * EPL + i * 8);
            if ((h % rep) == 0 && lane < D / 8) {
                // exactly one segment per kv head appends the new row to the cache (for FUTURE steps; this step uses the
                // locally rotated copy, so there is no intra-phase dependency on this write)
                uint4 kq;
                const float* s8 = sk + lane * 8;
                kq.x = pack_bf16(s8[0], s8[1]); kq.y = pack_bf16(s8[2], s8[3]);
                kq.z = pack_bf16(s8[4], s8[5]); kq.w = pack_bf16(s8[6], s8[7]);
                *reinterpret_cast<uint4*>(kc_l + ((size_t)hk * P.max_ctx + pos) * D + lane * 8) = kq;
                *reinterpret_cast<uint4*>(vc_l + ((size_t)hk * P.max_ctx + pos) * D + lane * 8) =
                    ll ? v8_ll((H + KVH + hk) * D + lane * 8) : ldcg4(vnew + lane * 8);
                __threadfence();                   // flags-in-data mode: ordered before this warp's partial packets (no barrier follows)
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
    const int cf0 = c * Lc, cf1 = min(cf0 + Lc, total);
    if (cf0 < cf1) {
        const int h_first = cf0 / ctx, h_last = (cf1 - 1) / ctx;
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
            const int c0 = (hh * ctx) / Lc;
            if (ll) {
                if constexpr (D <= 96) {
                    uint4* pk = P.att_ll + ((size_t)hh * P.att_maxp + (c - c0)) * 33;
                    float n1 = 0.f, n2 = 0.f;
                    if constexpr (D > 32) n1 = num[1];
                    if constexpr (D > 64) n2 = num[2];
                    st_packet4(pk + lane, num[0], n1, n2, out_flag);       // dims lane, lane + 32, lane + 64
                    if (lane == 0) st_packet4(pk + 32, M, L, 0.f, out_flag);
                }
                continue;
            }
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

template <int D, bool LL>
__global__ void __launch_bounds__(MG_THREADS, 1)
decode_mega_kernel(const MegaPlan* __restrict__ plan_g, int n_steps, long long* tokens_out, float* logits_out,
                   long long eos_id, long long pad_id, unsigned flag_base) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t s_full[MG_CONSUMERS * MG_SLOTS], s_empty[MG_CONSUMERS * MG_SLOTS];
    __shared__ float s_red[MG_CONSUMERS];
    __shared__ float s_rope[2 * 128];
    __shared__ unsigned s_epoch, s_prod_ord;
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
        s_prod_ord = 0;
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
        const int inflight = P.inflight, inflight_cur = P.inflight_cur;
        const int ppos0 = P.st->ctx_len;
        const unsigned bps = 5u * P.n_layers + 2u;     // grid barriers per step (consumer side)
        unsigned ord = 0;
        for (int stp = 0; stp < n_steps; ++stp)
            for (int i = 0; i < n_ops; ++i) {
                const MegaOp op = P.ops[i];
                // consumer epoch (grid barriers passed) while they work on this phase: 5 phases per layer, qkv first
                const int l = i >> 2, j = i & 3;
                const unsigned my_epoch = (unsigned)stp * bps + 5u * l + (j == 0 ? 0u : j + 1u);
                produce_phase(op, S, pc, inflight, inflight_cur, &s_epoch, my_epoch, lane, &s_prod_ord, ord);
                ord += (unsigned)cta_items(op);
                if (j == 0 && i + 1 < n_ops && !(P.ablate & 2)) {          // after qkv: this layer's K / V stream
                    const unsigned need = stp > 0 ? (unsigned)(stp - 1) * bps + 5u * l + 2u : 0u;
                    produce_attention<D>(P, l, ppos0 + stp, S, pc, inflight, inflight_cur, my_epoch + 1u, lane, &s_epoch, need);
                }
            }
        if (lane == 0) s_prod_ord = ord;
        return;
    }
    if (warp == MG_CONSUMERS + 1) {
        // ------------------------------------------------------------ L2 prefetcher: same item order, `pf_win` items ahead
        const int win = P.pf_win;
        if (win <= 0) return;
        const int ppos0 = P.st->ctx_len;
        unsigned ord = 0;
        for (int stp = 0; stp < n_steps; ++stp)
            for (int i = 0; i < n_ops; ++i) {
                const MegaOp op = P.ops[i];
                prefetch_phase(op, &s_prod_ord, ord, win, lane);
                ord += (unsigned)cta_items(op);
                if ((i & 3) == 0 && i + 1 < n_ops && !(P.ablate & 2))
                    prefetch_attention<D>(P, i >> 2, ppos0 + stp, &s_prod_ord, ord, win, lane);
            }
        return;
    }
    // ---------------------------------------------------------------- consumers
    unsigned epoch = 0, bars = 0;
    constexpr bool ll = LL;
    uint32_t cnt = 0;
    const int pos0 = P.st->ctx_len;                     // position == cache slot of the first token processed
    const int step0 = P.st->step;
    const bool tracing = P.trace != nullptr;
    const int ablate = P.ablate;                        // bring-up timing ablations (GVL_MEGA_ABLATE), 0 in production
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
        // phase ids of this step (flags-in-data mode): flag_base + step * bps + phase + 1, phase = 5 l + {qkv, attn, o, gate_up, down}
        const unsigned pid0 = flag_base + (unsigned)stp * (5u * P.n_layers + 2u) + 1u;
        auto boundary = [&]() {
            if (ll && !(ablate & 16)) phase_end(epoch, tid, &s_epoch);      // ablate 16: packets AND grid barriers (isolates the data path)
            else grid_barrier(P.grid_bar, epoch, bars, tid, &s_epoch, ablate);
        };
        for (int l = 0; l < P.n_layers; ++l) {
            const unsigned pq = pid0 + 5u * l;          // id of this layer's qkv phase
            // norm + qkv
            if (!(ablate & 4)) {
                if (ll && !(op.from_embed & 1)) stage_x_ll(op, pq - 1u, S, tid, warp, lane);           // x from down(l-1)
                else stage_x_vec(op, (op.from_embed & 1) ? emb_row : op.x, np, S, tid, warp, lane);
            }
            tr.mark(tid);
            long long* occ = tracing ? P.trace + (size_t)blockIdx.x * MEGA_TRACE_STRIDE + MEGA_TRACE_OCC_OFF : nullptr;
            gemv_items<LL>(op, P, emb_row, nullptr, S, cnt, tid, warp, lane, (occ && l == P.n_layers - 1) ? occ : nullptr, pq);
            op = P.ops[l * 4 + 1];
            tr.mark(tid); boundary(); tr.mark(tid);
            // rope + KV append + split-KV attention
            if (!(ablate & 2)) attention_phase<D, LL>(P, l, pos, S, cnt, warp, lane, pq, pq + 1u);
            tr.mark(tid);                               // keeps 3 marks per phase (no staging step here)
            tr.mark(tid); boundary(); tr.mark(tid);
            // merge + o_proj + residual
            if (!(ablate & 4)) {
                if (ll) stage_x_attn_ll(P, pq + 1u, S, pos + 1, tid);
                else stage_x_attn(P, S, pos + 1, tid);
            }
            tr.mark(tid);
            gemv_items<LL>(op, P, emb_row, nullptr, S, cnt, tid, warp, lane, (occ && l == P.n_layers - 1) ? occ + 8 : nullptr, pq + 2u);
            op = P.ops[l * 4 + 2];
            if (!ll) prefetch_norm(op, np, tid);
            tr.mark(tid); boundary(); tr.mark(tid);
            // norm + gate_up + SwiGLU
            if (!(ablate & 4)) {
                if (ll) stage_x_ll(op, pq + 2u, S, tid, warp, lane);
                else stage_x_vec(op, op.x, np, S, tid, warp, lane);
            }
            tr.mark(tid);
            gemv_items<LL>(op, P, emb_row, nullptr, S, cnt, tid, warp, lane, (occ && l == P.n_layers - 1) ? occ + 16 : nullptr, pq + 3u);
            op = P.ops[l * 4 + 3];
            tr.mark(tid); boundary(); tr.mark(tid);
            // down + residual
            if (!(ablate & 4)) {
                if (ll) stage_x_ll(op, pq + 3u, S, tid, warp, lane);
                else stage_x_vec(op, op.x, np, S, tid, warp, lane);
            }
            tr.mark(tid);
            gemv_items<LL>(op, P, emb_row, nullptr, S, cnt, tid, warp, lane, (occ && l == P.n_layers - 1) ? occ + 24 : nullptr, pq + 4u);
            op = P.ops[l * 4 + 4];
            if (!ll) prefetch_norm(op, np, tid);
            tr.mark(tid); boundary(); tr.mark(tid);
        }
        // norm + lm_head + bias + greedy pick
        if (ll) stage_x_ll(op, pid0 + 5u * P.n_layers - 1u, S, tid, warp, lane);
        else stage_x_vec(op, op.x, np, S, tid, warp, lane);
        tr.mark(tid);
        gemv_items<LL>(op, P, emb_row, logits_out ? logits_out + (size_t)step * P.vocab : nullptr, S, cnt, tid, warp, lane);
        tr.mark(tid); grid_barrier(P.grid_bar, epoch, bars, tid, &s_epoch, ablate); tr.mark(tid);
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
        if (stp + 1 < n_steps) grid_barrier(P.grid_bar, epoch, bars, tid, &s_epoch, ablate);   // the next step reads cur_token
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
    p->part_items = items;
    p->att_maxp = G / H + 2;
    const char* ab = getenv("GVL_MEGA_ABLATE");
    p->ablate = ab ? atoi(ab) : 0;
    const char* env = getenv("GVL_MEGA_INFLIGHT");
    p->inflight = env ? atoi(env) : 1;
    const char* ec = getenv("GVL_MEGA_INFLIGHT_CUR");
    p->inflight_cur = ec ? atoi(ec) : MG_SLOTS;
    const char* pw = getenv("GVL_MEGA_PFWIN");
    p->pf_win = pw ? atoi(pw) : 0;                     // measured slower with the prefetcher on (profiles/r1_decode.md)
    if (p->inflight < 1) p->inflight = 1;
    if (p->inflight > MG_SLOTS) p->inflight = MG_SLOTS;
    if (p->inflight_cur < p->inflight) p->inflight_cur = p->inflight;
    if (p->inflight_cur > MG_SLOTS) p->inflight_cur = MG_SLOTS;
    return (size_t)MG_RING_BYTES + p->x_bytes + (size_t)items * 32 <= (size_t)MG_SMEM_LIMIT;
}

size_t decode_mega_att_ws_bytes(const MegaPlan* p) {
    return (size_t)p->heads * p->att_maxp * (p->head_dim + 4) * sizeof(float);
}

unsigned decode_mega_phase_ids(const MegaPlan* p, int n_steps) { return (unsigned)n_steps * (5u * p->n_layers + 2u); }

// Flags-in-data mode needs every CTA in every phase's dependency chain (that is what bounds the skew between CTAs to one phase
// and makes the in-place reuse of the x / qkv / mid / partial buffers safe without grid barriers): every GEMV op of a layer must
// have at least gridDim units, outputs in whole packets; the partial packets hold 3 floats per lane -> head_dim <= 96, and the
// merge keeps <= 8 partials per head in registers.
bool decode_mega_ll_supported(const MegaPlan* p) {
    const int G = num_sms();
    if (p->head_dim > 96 || p->att_maxp > 8) return false;
    for (int i = 0; i < p->n_layers * 4; ++i) {
        const MegaOp& op = p->ops[i];
        if (op.units < G || op.n_out % 8 != 0 || op.K % 4 != 0 || op.K > 8192) return false;
    }
    return p->ops[p->n_layers * 4].K <= 8192;
}

size_t decode_mega_att_ll_packets(const MegaPlan* p) { return (size_t)p->heads * p->att_maxp * 33; }

int decode_mega_launch(const MegaPlan* hp, const MegaPlan* plan_dev, int n_steps, long long* tokens_out, float* logits_out,
                       long long eos_id, long long pad_id, unsigned flag_base, cudaStream_t s) {
    const size_t smem = (size_t)MG_RING_BYTES + hp->x_bytes + (size_t)hp->part_items * 32;
    using KernelT = void (*)(const MegaPlan*, int, long long*, float*, long long, long long, unsigned);
    const int di = hp->head_dim == 64 ? 0 : hp->head_dim == 96 ? 1 : 2;
    KernelT kern = hp->use_ll ? (di == 0 ? decode_mega_kernel<64, true> : decode_mega_kernel<96, true>)
                              : (di == 0 ? decode_mega_kernel<64, false> : di == 1 ? decode_mega_kernel<96, false>
                                                                                 : decode_mega_kernel<128, false>);
    static size_t attr_set_tab[64][2][3] = {};         // per device (cudaFuncSetAttribute is per device), per kernel variant
    int dev = 0;
    cudaGetDevice(&dev);
    size_t* attr_set = attr_set_tab[dev & 63][hp->use_ll ? 1 : 0];
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
    if (cudaLaunchKernelEx(&cfg, kern, plan_dev, n_steps, tokens_out, logits_out, eos_id, pad_id, flag_base) != cudaSuccess)
        return GVL_ERR_CUDA;
    g_launch_count++;
    return cudaGetLastError() == cudaSuccess ? GVL_OK : GVL_ERR_CUDA;
}

}  // namespace gvl
