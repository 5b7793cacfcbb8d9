// One persistent kernel per decode step ("megakernel"): the 4 x L + 1 weight matrices of a step are streamed
// through a shared-memory ring by ONE producer warp per SM that never stops for phase boundaries (weights do not
// depend on activations), while 8 consumer warps per SM walk the phases of the step
//     embed | per layer: [norm+qkv] B [rope + KV append + split-KV attention + combine] B [o_proj+res] B
//           [norm+gate_up+SwiGLU] B [down+res] B | [norm+lm_head+bias] B argmax/bookkeeping
// separated by grid-wide barriers B (atomic counter, all 148 CTAs co-resident, cooperative launch).
// While the consumers sit in a barrier the ring keeps filling (~190 KB per SM, ~4 us of HBM time chip-wide), so the
// HBM stream does not drain at kernel/phase boundaries -- the ~5 us fixed cost per GEMV launch measured for the
// per-op kernels (profiles/r1_decode.md) disappears.
//
// Reference semantics per phase: Phi3DecoderLayer / LlamaDecoderLayer with q_len = 1 (modeling_phi3.py:1034-1095,
// 629-775, 413-445; modeling_llama.py:699-760), lm_head + .float() (modeling_phi3.py:1525-1526), greedy pick of
// HF generate (llava_next_video.py:655-661). Rounding points identical to gemv.cu / decode.cu.
#include "gvl_internal.h"
#include "ptx.cuh"
#include "decode.h"
#include "decode_mega.h"

namespace gvl {

namespace {

constexpr int MG_CONSUMERS = 8;
constexpr int MG_THREADS = 32 * (MG_CONSUMERS + 1);
constexpr int MG_SLOT_BYTES = 8192;                       // one segment (<= 4096 bf16)
constexpr int MG_STAGE_BYTES = MG_CONSUMERS * MG_SLOT_BYTES;
constexpr int MG_STAGES = 3;
constexpr int MG_XMAX = 14336;                            // largest K staged (elements; Llama-3-8B ffn)
constexpr int MG_SMEM = MG_STAGES * MG_STAGE_BYTES + MG_XMAX * 2 + 1024;
constexpr int MG_SPLIT = 128;                             // context tokens per attention task

__device__ __forceinline__ float wsum_m(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float dot8m(uint4 w, uint4 x) {
    float2 a, b;
    float s;
    a = unpack_bf16(w.x); b = unpack_bf16(x.x); s = a.x * b.x + a.y * b.y;
    a = unpack_bf16(w.y); b = unpack_bf16(x.y); s += a.x * b.x + a.y * b.y;
    a = unpack_bf16(w.z); b = unpack_bf16(x.z); s += a.x * b.x + a.y * b.y;
    a = unpack_bf16(w.w); b = unpack_bf16(x.w); s += a.x * b.x + a.y * b.y;
    return s;
}
__device__ __forceinline__ void bulk_g2s_m(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint4 ldcg4(const void* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ uint4 ldg_stream4(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// consumers-only grid barrier (named barrier 1 = the 256 consumer threads of this CTA)
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& epoch, int tid) {
    __threadfence();
    asm volatile("bar.sync 1, 256;" ::: "memory");
    ++epoch;
    if (tid == 0) {
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
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
}

// optional per-CTA phase trace (clock64 at: x staged / work done / barrier passed), MegaPlan::trace != nullptr
struct Tracer {
    long long* p;
    __device__ __forceinline__ void mark(int tid) {
        if (p != nullptr && tid == 0) *p++ = clock64();
    }
};

__device__ __forceinline__ int units_of(int gl, int n_units, int TW) { return gl < n_units ? (n_units - gl + TW - 1) / TW : 0; }
__device__ __forceinline__ int row_of(const MegaOp& op, int unit, int sel) {
    return op.act == 3 ? (unit / 128) * 256 + (unit % 128) + sel * 128 : unit;
}

// ------------------------------------------------------------------ consumer side of one GEMV phase
struct RingState {
    int stage;
    uint32_t phase;
};

__device__ __forceinline__ void gemv_phase(const MegaOp& op_g, __nv_bfloat16* sx, uint8_t* ring, uint64_t* s_full,
                                           uint64_t* s_empty, float* s_red, RingState& rs, int tid, int warp, int lane,
                                           Tracer& tr) {
    const MegaOp op = op_g;      // registers: the plan lives in global memory and the stores below could alias it
    const int K = op.K, kv = K / 8;
    // ---- stage x (global, written by other CTAs in the previous phase -> L1-bypassing loads), optional RMSNorm
    float ss = 0.f;
    for (int i = tid; i < kv; i += 256) {
        const uint4 v = ldcg4(op.x + (size_t)i * 8);
        reinterpret_cast<uint4*>(sx)[i] = v;
        if (op.norm_w) {
            float2 f;
            f = unpack_bf16(v.x); ss += f.x * f.x + f.y * f.y;
            f = unpack_bf16(v.y); ss += f.x * f.x + f.y * f.y;
            f = unpack_bf16(v.z); ss += f.x * f.x + f.y * f.y;
            f = unpack_bf16(v.w); ss += f.x * f.x + f.y * f.y;
        }
    }
    if (op.norm_w) {
        ss = wsum_m(ss);
        if (lane == 0) s_red[warp] = ss;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < MG_CONSUMERS; ++w) t += s_red[w];
        const float rstd = rsqrtf(t / K + op.eps);
        for (int i = tid; i < kv; i += 256) {
            const uint4 wv = __ldg(reinterpret_cast<const uint4*>(op.norm_w) + i);
            uint4 v = reinterpret_cast<uint4*>(sx)[i], o;
            float2 f, g;
            f = unpack_bf16(v.x); g = unpack_bf16(wv.x); o.x = pack_bf16(bf16r(f.x * rstd) * g.x, bf16r(f.y * rstd) * g.y);
            f = unpack_bf16(v.y); g = unpack_bf16(wv.y); o.y = pack_bf16(bf16r(f.x * rstd) * g.x, bf16r(f.y * rstd) * g.y);
            f = unpack_bf16(v.z); g = unpack_bf16(wv.z); o.z = pack_bf16(bf16r(f.x * rstd) * g.x, bf16r(f.y * rstd) * g.y);
            f = unpack_bf16(v.w); g = unpack_bf16(wv.w); o.w = pack_bf16(bf16r(f.x * rstd) * g.x, bf16r(f.y * rstd) * g.y);
            reinterpret_cast<uint4*>(sx)[i] = o;
        }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    tr.mark(tid);

    const int TW = gridDim.x * MG_CONSUMERS;
    const int nsel = op.act == 3 ? 2 : 1;
    const int ipu = nsel * op.nseg;
    const int gl = blockIdx.x * MG_CONSUMERS + warp;
    const int my_items = units_of(gl, op.units, TW) * ipu;
    const int max_items = units_of(blockIdx.x * MG_CONSUMERS, op.units, TW) * ipu;   // slot 0 always has the most
    const int chunks = op.seg_len / 8;
    float acc0 = 0.f, acc1 = 0.f;
    int unit = gl, sel = 0, seg = 0;
    for (int i = 0; i < max_items; ++i) {
        ptx::mbar_wait(ptx::smem_u32(&s_full[rs.stage]), rs.phase);
        if (i < my_items) {
            const uint4* wseg = reinterpret_cast<const uint4*>(ring + (size_t)rs.stage * MG_STAGE_BYTES + (size_t)warp * MG_SLOT_BYTES);
            const uint4* xs = reinterpret_cast<const uint4*>(sx + (size_t)seg * op.seg_len);
            float part = 0.f;
#pragma unroll 4
            for (int c = lane; c < chunks; c += 32) part += dot8m(wseg[c], xs[c]);
            if (sel == 0) acc0 += part; else acc1 += part;
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&s_empty[rs.stage]));
        if (++rs.stage == MG_STAGES) { rs.stage = 0; rs.phase ^= 1; }
        if (i < my_items) {
            if (++seg == op.nseg) {
                seg = 0;
                if (++sel == nsel) {
                    // ---- unit finished
                    sel = 0;
                    acc0 = wsum_m(acc0);
                    if (op.act == 3) acc1 = wsum_m(acc1);
                    if (lane == 0) {
                        if (op.act == 3) {
                            const float g = bf16r(acc0), u = bf16r(acc1);
                            reinterpret_cast<__nv_bfloat16*>(op.out)[unit] = __float2bfloat16_rn(u * bf16r(silu_f(g)));
                        } else {
                            float y = acc0;
                            if (op.bias) y += __bfloat162float(op.bias[unit]);
                            y = bf16r(y);
                            if (op.residual) {
                                // written by another CTA in an earlier phase -> L1-bypassing load
                                const unsigned short raw = __ldcg(reinterpret_cast<const unsigned short*>(op.residual) + unit);
                                y = bf16r(y + __uint_as_float((uint32_t)raw << 16));
                            }
                            if (op.out_f32) reinterpret_cast<float*>(op.out)[unit] = y;
                            else reinterpret_cast<__nv_bfloat16*>(op.out)[unit] = __float2bfloat16_rn(y);
                        }
                    }
                    acc0 = acc1 = 0.f;
                    unit += TW;
                }
            }
        }
    }
}

// ------------------------------------------------------------------ producer side of one GEMV phase (lanes 0..7)
__device__ __forceinline__ void produce_phase(const MegaOp& op_g, uint32_t ring_u32, uint64_t* s_full, uint64_t* s_empty,
                                              RingState& rs, int lane) {
    const MegaOp op = op_g;
    const int TW = gridDim.x * MG_CONSUMERS;
    const int nsel = op.act == 3 ? 2 : 1;
    const int ipu = nsel * op.nseg;
    const int gl = blockIdx.x * MG_CONSUMERS + lane;                 // lane w feeds consumer warp w's slot
    const int my_items = lane < MG_CONSUMERS ? units_of(gl, op.units, TW) * ipu : 0;
    const int max_items = units_of(blockIdx.x * MG_CONSUMERS, op.units, TW) * ipu;
    const uint32_t seg_bytes = (uint32_t)op.seg_len * 2;
    int unit = gl, sel = 0, seg = 0;
    for (int i = 0; i < max_items; ++i) {
        ptx::mbar_wait(ptx::smem_u32(&s_empty[rs.stage]), rs.phase ^ 1);
        const bool valid = i < my_items;
        const unsigned mask = __ballot_sync(0xffffffffu, valid);
        const uint32_t bar = ptx::smem_u32(&s_full[rs.stage]);
        if (lane == 0) ptx::mbar_arrive_expect_tx(bar, __popc(mask) * seg_bytes);
        __syncwarp();
        if (valid) {
            const int row = row_of(op, unit, sel);
            bulk_g2s_m(ring_u32 + rs.stage * MG_STAGE_BYTES + lane * MG_SLOT_BYTES,
                       op.W + (size_t)row * op.ldw + (size_t)seg * op.seg_len, seg_bytes, bar);
            if (++seg == op.nseg) { seg = 0; if (++sel == nsel) { sel = 0; unit += TW; } }
        }
        if (++rs.stage == MG_STAGES) { rs.stage = 0; rs.phase ^= 1; }
    }
}

// ------------------------------------------------------------------ attention phase (warp task = (head, 128-token split))
template <int D>
__device__ __forceinline__ void attention_phase(const MegaPlan& P, int layer, float* s_warp /* [2*D] floats per warp */,
                                                int warp, int lane) {
    constexpr int EPL = D / 4, VPL = EPL / 8, half = D / 2;
    const int H = P.heads, KVH = P.kv_heads;
    const int pos = P.st->ctx_len;                    // position == cache slot of the token being processed
    const int ctx = pos + 1;
    const int n_act = (ctx + MG_SPLIT - 1) / MG_SPLIT;
    const int n_tasks = H * n_act;
    const int TW = gridDim.x * MG_CONSUMERS;
    const __nv_bfloat16* qkv = P.qkv;
    __nv_bfloat16* kc_l = P.kv + (size_t)layer * 2 * KVH * P.max_ctx * D;
    __nv_bfloat16* vc_l = kc_l + (size_t)KVH * P.max_ctx * D;
    const __nv_bfloat16* cp = P.rope_cos + (size_t)pos * D;
    const __nv_bfloat16* sp = P.rope_sin + (size_t)pos * D;
    float* sq = s_warp;          // rotated q   [D]
    float* sk = s_warp + D;      // rotated new k [D]
    const int sub = lane & 3, tg = lane >> 2;         // 4 lanes per token, 8 tokens per pass
    const int nsplit_ws = (P.max_ctx + MG_SPLIT - 1) / MG_SPLIT;
    for (int task = blockIdx.x * MG_CONSUMERS + warp; task < n_tasks; task += TW) {
        const int h = task / n_act, s = task % n_act;
        const int hk = h / (H / KVH);
        const int t0 = s * MG_SPLIT, t1 = min(t0 + MG_SPLIT, ctx);
        // ---- RoPE of q (and of the new k) for this head: q_embed = bf16(bf16(q*cos) + bf16(rot(q)*sin))
        __syncwarp();
        for (int j = lane; j < half; j += 32) {
            const float c1 = __bfloat162float(cp[j]), s1 = __bfloat162float(sp[j]);
            const float c2 = __bfloat162float(cp[j + half]), s2 = __bfloat162float(sp[j + half]);
            const __nv_bfloat16* qh = qkv + (size_t)h * D;
            const float q1 = __bfloat162float(__ldcg(qh + j)), q2 = __bfloat162float(__ldcg(qh + j + half));
            sq[j] = bf16r(bf16r(q1 * c1) + bf16r(-q2 * s1));
            sq[j + half] = bf16r(bf16r(q2 * c2) + bf16r(q1 * s2));
            const __nv_bfloat16* kh = qkv + (size_t)(H + hk) * D;
            const float k1 = __bfloat162float(__ldcg(kh + j)), k2 = __bfloat162float(__ldcg(kh + j + half));
            sk[j] = bf16r(bf16r(k1 * c1) + bf16r(-k2 * s1));
            sk[j + half] = bf16r(bf16r(k2 * c2) + bf16r(k1 * s2));
        }
        __syncwarp();
        const bool has_new = pos >= t0 && pos < t1;
        const __nv_bfloat16* vnew = qkv + (size_t)(H + KVH + hk) * D;
        if (has_new && (h % (H / KVH)) == 0) {
            // exactly one task per kv head appends the new row to the cache (for FUTURE steps; this step's tasks use the
            // locally rotated copy, so there is no intra-phase dependency on this write)
            for (int j = lane; j < D; j += 32) {
                kc_l[((size_t)hk * P.max_ctx + pos) * D + j] = __float2bfloat16_rn(sk[j]);
                vc_l[((size_t)hk * P.max_ctx + pos) * D + j] = __ldcg(vnew + j);
            }
        }
        float qf[EPL];
#pragma unroll
        for (int i = 0; i < EPL; ++i) qf[i] = sq[sub * EPL + i];
        const __nv_bfloat16* kbase = kc_l + (size_t)hk * P.max_ctx * D;
        const __nv_bfloat16* vbase = vc_l + (size_t)hk * P.max_ctx * D;
        // ---- scores: 16 passes of 8 tokens, K loads of 4 passes in flight
        constexpr int NP = MG_SPLIT / 8;
        float sc[NP];
        float lmax = -INFINITY;
#pragma unroll
        for (int pb = 0; pb < NP; pb += 4) {
            uint4 kr[4][VPL];
#pragma unroll
            for (int pp = 0; pp < 4; ++pp) {
                const int t = t0 + (pb + pp) * 8 + tg;
                const bool ok = t < t1 && t != pos;
                const __nv_bfloat16* src = kbase + (size_t)(ok ? t : t0) * D + sub * EPL;
#pragma unroll
                for (int i = 0; i < VPL; ++i) kr[pp][i] = ldg_stream4(reinterpret_cast<const uint4*>(src) + i);
            }
#pragma unroll
            for (int pp = 0; pp < 4; ++pp) {
                const int t = t0 + (pb + pp) * 8 + tg;
                float sv = 0.f;
                if (t == pos) {
#pragma unroll
                    for (int i = 0; i < EPL; ++i) sv += qf[i] * sk[sub * EPL + i];
                } else {
#pragma unroll
                    for (int i = 0; i < VPL; ++i) {
                        const uint4 v = kr[pp][i];
                        float2 f;
                        f = unpack_bf16(v.x); sv += qf[i * 8 + 0] * f.x + qf[i * 8 + 1] * f.y;
                        f = unpack_bf16(v.y); sv += qf[i * 8 + 2] * f.x + qf[i * 8 + 3] * f.y;
                        f = unpack_bf16(v.z); sv += qf[i * 8 + 4] * f.x + qf[i * 8 + 5] * f.y;
                        f = unpack_bf16(v.w); sv += qf[i * 8 + 6] * f.x + qf[i * 8 + 7] * f.y;
                    }
                }
                sv += __shfl_xor_sync(0xffffffffu, sv, 1);
                sv += __shfl_xor_sync(0xffffffffu, sv, 2);
                sv = (t < t1) ? sv * P.scale : -INFINITY;
                sc[pb + pp] = sv;
                lmax = fmaxf(lmax, sv);
            }
        }
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
        // ---- p = exp(s - max) rounded to bf16; o += p * v
        float o[EPL];
#pragma unroll
        for (int i = 0; i < EPL; ++i) o[i] = 0.f;
        float lsum = 0.f;
#pragma unroll
        for (int pb = 0; pb < NP; pb += 4) {
            uint4 vr[4][VPL];
#pragma unroll
            for (int pp = 0; pp < 4; ++pp) {
                const int t = t0 + (pb + pp) * 8 + tg;
                const bool ok = t < t1 && t != pos;
                const __nv_bfloat16* src = vbase + (size_t)(ok ? t : t0) * D + sub * EPL;
#pragma unroll
                for (int i = 0; i < VPL; ++i) vr[pp][i] = ldg_stream4(reinterpret_cast<const uint4*>(src) + i);
            }
#pragma unroll
            for (int pp = 0; pp < 4; ++pp) {
                const int t = t0 + (pb + pp) * 8 + tg;
                const float p = bf16r(__expf(sc[pb + pp] - lmax));
                lsum += p;
                if (t == pos) {
#pragma unroll
                    for (int i = 0; i < EPL; ++i) o[i] += p * __bfloat162float(__ldcg(vnew + sub * EPL + i));
                } else {
#pragma unroll
                    for (int i = 0; i < VPL; ++i) {
                        const uint4 v = vr[pp][i];
                        float2 f;
                        f = unpack_bf16(v.x); o[i * 8 + 0] += p * f.x; o[i * 8 + 1] += p * f.y;
                        f = unpack_bf16(v.y); o[i * 8 + 2] += p * f.x; o[i * 8 + 3] += p * f.y;
                        f = unpack_bf16(v.z); o[i * 8 + 4] += p * f.x; o[i * 8 + 5] += p * f.y;
                        f = unpack_bf16(v.w); o[i * 8 + 6] += p * f.x; o[i * 8 + 7] += p * f.y;
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < EPL; ++i) {
            o[i] += __shfl_xor_sync(0xffffffffu, o[i], 4);
            o[i] += __shfl_xor_sync(0xffffffffu, o[i], 8);
            o[i] += __shfl_xor_sync(0xffffffffu, o[i], 16);
        }
        if (sub != 0) lsum = 0.f;
        lsum = wsum_m(lsum);
        float* wrow = P.att_ws + ((size_t)h * nsplit_ws + s) * (D + 2);
        if (lane < 4) {
#pragma unroll
            for (int i = 0; i < EPL; ++i) wrow[2 + lane * EPL + i] = o[i];
        }
        if (lane == 0) { wrow[0] = lmax; wrow[1] = lsum; }
        // ---- last-arriving split of head h merges
        __threadfence();
        __syncwarp();
        int last = 0;
        if (lane == 0) last = (atomicAdd(&P.att_counters[h], 1) == n_act - 1) ? 1 : 0;
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {
            __threadfence();
            const float* base = P.att_ws + (size_t)h * nsplit_ws * (D + 2);
            float M = -INFINITY;
            for (int ss_ = 0; ss_ < n_act; ++ss_) M = fmaxf(M, __ldcg(base + (size_t)ss_ * (D + 2)));
            for (int d = lane; d < D; d += 32) {
                float num = 0.f, den = 0.f;
                for (int ss_ = 0; ss_ < n_act; ++ss_) {
                    const float* r = base + (size_t)ss_ * (D + 2);
                    const float wgt = __expf(__ldcg(r) - M);
                    num += wgt * __ldcg(r + 2 + d);
                    den += wgt * __ldcg(r + 1);
                }
                P.attn_out[(size_t)h * D + d] = __float2bfloat16_rn(den > 0.f ? num / den : 0.f);
            }
            if (lane == 0) P.att_counters[h] = 0;
        }
    }
}

__global__ void __launch_bounds__(MG_THREADS, 1)
decode_mega_kernel(const MegaPlan* __restrict__ plan_g, long long* tokens_out, float* logits_out, long long eos_id,
                   long long pad_id) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* ring = smem;
    __nv_bfloat16* sx = reinterpret_cast<__nv_bfloat16*>(smem + MG_STAGES * MG_STAGE_BYTES);
    __shared__ __align__(8) uint64_t s_full[MG_STAGES], s_empty[MG_STAGES];
    __shared__ float s_red[MG_CONSUMERS];
    // the attention phase never overlaps a GEMV phase of the same CTA: its per-warp scratch aliases the x staging area
    float (*s_att)[2 * 128] = reinterpret_cast<float (*)[2 * 128]>(sx);
    const MegaPlan& P = *plan_g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < MG_STAGES; ++s) {
            ptx::mbar_init(ptx::smem_u32(&s_full[s]), 1);
            ptx::mbar_init(ptx::smem_u32(&s_empty[s]), MG_CONSUMERS);
        }
        ptx::fence_mbar_init();
    }
    __syncthreads();
    RingState rs{0, 0};
    const int n_ops = P.n_layers * 4 + 1;

    if (warp == MG_CONSUMERS) {
        // ------------------------------------------------------------ producer: the whole step's weights, in order
        const uint32_t ring_u32 = ptx::smem_u32(ring);
        for (int i = 0; i < n_ops; ++i) produce_phase(P.ops[i], ring_u32, s_full, s_empty, rs, lane);
        return;
    }
    // ---------------------------------------------------------------- consumers
    unsigned epoch = 0;
    Tracer tr{P.trace ? P.trace + (size_t)blockIdx.x * MEGA_TRACE_STRIDE : nullptr};
    tr.mark(tid);
    // embed the current token into the residual stream (CTA 0), everyone waits
    if (blockIdx.x == 0) {
        const long long tok = P.st->cur_token;
        const uint4* src = reinterpret_cast<const uint4*>(P.embed + (size_t)tok * P.dim);
        for (int i = tid; i < P.dim / 8; i += 256) reinterpret_cast<uint4*>(P.x)[i] = src[i];
    }
    tr.mark(tid); grid_barrier(P.grid_bar, epoch, tid); tr.mark(tid);
    for (int l = 0; l < P.n_layers; ++l) {
        gemv_phase(P.ops[l * 4 + 0], sx, ring, s_full, s_empty, s_red, rs, tid, warp, lane, tr);     // norm + qkv
        tr.mark(tid); grid_barrier(P.grid_bar, epoch, tid); tr.mark(tid);
        if (P.head_dim == 96) attention_phase<96>(P, l, s_att[warp], warp, lane);
        else if (P.head_dim == 128) attention_phase<128>(P, l, s_att[warp], warp, lane);
        else attention_phase<64>(P, l, s_att[warp], warp, lane);
        tr.mark(tid);                                   // keeps 3 marks per phase (no staging step here)
        tr.mark(tid); grid_barrier(P.grid_bar, epoch, tid); tr.mark(tid);
        gemv_phase(P.ops[l * 4 + 1], sx, ring, s_full, s_empty, s_red, rs, tid, warp, lane, tr);     // o_proj + residual
        tr.mark(tid); grid_barrier(P.grid_bar, epoch, tid); tr.mark(tid);
        gemv_phase(P.ops[l * 4 + 2], sx, ring, s_full, s_empty, s_red, rs, tid, warp, lane, tr);     // norm + gate_up + SwiGLU
        tr.mark(tid); grid_barrier(P.grid_bar, epoch, tid); tr.mark(tid);
        gemv_phase(P.ops[l * 4 + 3], sx, ring, s_full, s_empty, s_red, rs, tid, warp, lane, tr);     // down + residual
        tr.mark(tid); grid_barrier(P.grid_bar, epoch, tid); tr.mark(tid);
    }
    gemv_phase(P.ops[P.n_layers * 4], sx, ring, s_full, s_empty, s_red, rs, tid, warp, lane, tr);    // norm + lm_head + bias
    tr.mark(tid); grid_barrier(P.grid_bar, epoch, tid); tr.mark(tid);
    // ---- greedy pick + bookkeeping (CTA 0): tokens_out[step] = argmax (or pad after EOS), ctx_len++, step++
    if (blockIdx.x == 0) {
        __shared__ float sv_[MG_CONSUMERS];
        __shared__ int si_[MG_CONSUMERS];
        DecodeState* st = P.st;
        const int step = st->step;
        float best = -INFINITY;
        int bi = 0x7fffffff;
        for (int i = tid; i < P.vocab; i += 256) {
            const float v = __ldcg(P.logits + i);
            if (logits_out) logits_out[(size_t)step * P.vocab + i] = v;
            if (v > best || (v == best && i < bi)) { best = v; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (lane == 0) { sv_[warp] = best; si_[warp] = bi; }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (tid == 0) {
            for (int w = 1; w < MG_CONSUMERS; ++w)
                if (sv_[w] > best || (sv_[w] == best && si_[w] < bi)) { best = sv_[w]; bi = si_[w]; }
            long long tok = bi;
            if (st->finished) tok = pad_id;
            else if (eos_id >= 0 && tok == eos_id) st->finished = 1;
            if (tokens_out) tokens_out[step] = tok;
            st->cur_token = tok;
            st->ctx_len = st->ctx_len + 1;
            st->attn_len = st->ctx_len;
            st->step = step + 1;
        }
    }
}

}  // namespace

size_t decode_mega_smem() { return MG_SMEM; }

int decode_mega_launch(const MegaPlan* plan_dev, unsigned* grid_bar, long long* tokens_out, float* logits_out,
                       long long eos_id, long long pad_id, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(decode_mega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MG_SMEM) != cudaSuccess)
            return GVL_ERR_CUDA;
        attr_set = true;
    }
    if (cudaMemsetAsync(grid_bar, 0, sizeof(unsigned), s) != cudaSuccess) return GVL_ERR_CUDA;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(num_sms());
    cfg.blockDim = dim3(MG_THREADS);
    cfg.dynamicSmemBytes = MG_SMEM;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, decode_mega_kernel, plan_dev, tokens_out, logits_out, eos_id, pad_id) != cudaSuccess)
        return GVL_ERR_CUDA;
    g_launch_count++;
    return cudaGetLastError() == cudaSuccess ? GVL_OK : GVL_ERR_CUDA;
}

}  // namespace gvl
