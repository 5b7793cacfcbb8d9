// Decode-side kernels (q_len = 1): everything here is HBM-bound weight / KV streaming, so the
// design rules are coalesced 128-bit loads, many bytes in flight per SM and no tensor cores.
//   (the GEMV lives in gemv.cu)
//   decode_attn_kernel     one query vs the KV cache, split along the context (flash-decoding style)
//   argmax / token kernels greedy sampling + device-side step bookkeeping (so a step is graph-replayable)
// Reference call sites: Phi3DecoderLayer.forward with q_len=1 (modeling_phi3.py:1034-1095, 629-775),
// lm_head + .float() (modeling_phi3.py:1525-1526), HF greedy loop (llava_next_video.py:655-661).
#include "gvl_internal.h"
#include "ptx.cuh"
#include "decode.h"

namespace gvl {

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}


// ---------------------------------------------------------------- decode attention
constexpr int DA_THREADS = 128;
constexpr int DA_CHUNK = 128;  // context tokens per CTA

// Last-arriving split of a head merges the per-split (max, sum, partial-output) triples (flash-decoding combine),
// so the whole single-query attention is ONE launch. counters[h] is self-resetting.
template <int D>
__device__ __forceinline__ void da_finish(float* ws, int* counters, __nv_bfloat16* o_out, int h, int nsplit, int tid) {
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&counters[h], 1) == nsplit - 1) ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const float* base = ws + (size_t)h * nsplit * (D + 2);
    if (tid < D) {
        float M = -INFINITY;
        for (int s = 0; s < nsplit; ++s) M = fmaxf(M, __ldcg(base + (size_t)s * (D + 2)));
        float num = 0.f, den = 0.f;
        for (int s = 0; s < nsplit; ++s) {
            const float* r = base + (size_t)s * (D + 2);
            const float m = __ldcg(r);
            if (m == -INFINITY) continue;
            const float wgt = __expf(m - M);
            num += wgt * __ldcg(r + 2 + tid);
            den += wgt * __ldcg(r + 1);
        }
        o_out[(size_t)h * D + tid] = __float2bfloat16_rn(den > 0.f ? num / den : 0.f);
    }
    if (tid == 0) counters[h] = 0;
}

// grid (heads, max_splits). 4 lanes share one token (D/4 elements each); 32 tokens per CTA iteration.
template <int D>
__device__ __forceinline__ void
decode_attn_body(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ kc,
                 const __nv_bfloat16* __restrict__ vc, float* ws, const int* __restrict__ ctx_len_dev,
                 int heads, int kv_heads, int max_ctx, float scale, int* counters, __nv_bfloat16* o_out, int split, int nsplit) {
    constexpr int EPL = D / 4;  // elements per lane
    constexpr int VPL = EPL / 8;
    const int ctx = *ctx_len_dev;  // tokens in the cache INCLUDING the one appended this step
    const int h = blockIdx.x;
    const int hk = h / (heads / kv_heads);
    const int t0 = split * DA_CHUNK;
    float* wrow = ws + ((size_t)h * nsplit + split) * (D + 2);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (t0 >= ctx) {
        if (tid == 0) { wrow[0] = -INFINITY; wrow[1] = 0.f; }
        da_finish<D>(ws, counters, o_out, h, nsplit, tid);
        return;
    }
    const int t1 = min(t0 + DA_CHUNK, ctx);
    __shared__ float s_red[DA_THREADS / 32];
    __shared__ float s_o[DA_THREADS / 32][D];

    const int sub = lane & 3;         // which quarter of the head dim
    const int tl = tid >> 2;          // token lane within the 32-token group
    float qf[EPL];
    {
        const uint4* qp = reinterpret_cast<const uint4*>(q + (size_t)h * D + sub * EPL);
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            uint4 v = __ldg(qp + i);
            float2 f;
            f = unpack_bf16(v.x); qf[i * 8 + 0] = f.x; qf[i * 8 + 1] = f.y;
            f = unpack_bf16(v.y); qf[i * 8 + 2] = f.x; qf[i * 8 + 3] = f.y;
            f = unpack_bf16(v.z); qf[i * 8 + 4] = f.x; qf[i * 8 + 5] = f.y;
            f = unpack_bf16(v.w); qf[i * 8 + 6] = f.x; qf[i * 8 + 7] = f.y;
        }
    }
    const __nv_bfloat16* kbase = kc + (size_t)hk * max_ctx * D;
    const __nv_bfloat16* vbase = vc + (size_t)hk * max_ctx * D;
    // ---- issue EVERY K and V load of this thread's tokens up front (one DRAM round trip instead of eight)
    constexpr int NIT = DA_CHUNK / (DA_THREADS / 4);
    uint4 kreg[NIT][VPL], vreg[NIT][VPL];
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
        const int t = t0 + tl + it * (DA_THREADS / 4);
        const bool ok = t < t1;
        const size_t off = (size_t)(ok ? t : t0) * D + sub * EPL;
#pragma unroll
        for (int i = 0; i < VPL; ++i) kreg[it][i] = ldg_stream(reinterpret_cast<const uint4*>(kbase + off) + i);
    }
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
        const int t = t0 + tl + it * (DA_THREADS / 4);
        const bool ok = t < t1;
        const size_t off = (size_t)(ok ? t : t0) * D + sub * EPL;
#pragma unroll
        for (int i = 0; i < VPL; ++i) vreg[it][i] = ldg_stream(reinterpret_cast<const uint4*>(vbase + off) + i);
    }
    // ---- pass 1: scores
    float lmax = -INFINITY;
    float sc[NIT];
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
        const int t = t0 + tl + it * (DA_THREADS / 4);
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const uint4 v = kreg[it][i];
            float2 f;
            f = unpack_bf16(v.x); s += qf[i * 8 + 0] * f.x + qf[i * 8 + 1] * f.y;
            f = unpack_bf16(v.y); s += qf[i * 8 + 2] * f.x + qf[i * 8 + 3] * f.y;
            f = unpack_bf16(v.z); s += qf[i * 8 + 4] * f.x + qf[i * 8 + 5] * f.y;
            f = unpack_bf16(v.w); s += qf[i * 8 + 6] * f.x + qf[i * 8 + 7] * f.y;
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s = (t < t1) ? s * scale : -INFINITY;
        sc[it] = s;
        lmax = fmaxf(lmax, s);
    }
    lmax = warp_max(lmax);
    if (lane == 0) s_red[warp] = lmax;
    __syncthreads();
    float bmax = s_red[0];
#pragma unroll
    for (int w = 1; w < DA_THREADS / 32; ++w) bmax = fmaxf(bmax, s_red[w]);
    // ---- pass 2: p = exp(s - max) (rounded to bf16 like the prefill kernel), o += p * v
    float o[EPL];
#pragma unroll
    for (int i = 0; i < EPL; ++i) o[i] = 0.f;
    float lsum = 0.f;
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
        const float p = bf16r(__expf(sc[it] - bmax));    // exp(-inf) = 0 for the padded tokens
        lsum += p;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const uint4 v = vreg[it][i];
            float2 f;
            f = unpack_bf16(v.x); o[i * 8 + 0] += p * f.x; o[i * 8 + 1] += p * f.y;
            f = unpack_bf16(v.y); o[i * 8 + 2] += p * f.x; o[i * 8 + 3] += p * f.y;
            f = unpack_bf16(v.z); o[i * 8 + 4] += p * f.x; o[i * 8 + 5] += p * f.y;
            f = unpack_bf16(v.w); o[i * 8 + 6] += p * f.x; o[i * 8 + 7] += p * f.y;
        }
    }
    // reduce over the 8 token lanes of the warp (lanes with equal `sub`), then over warps
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
        o[i] += __shfl_xor_sync(0xffffffffu, o[i], 4);
        o[i] += __shfl_xor_sync(0xffffffffu, o[i], 8);
        o[i] += __shfl_xor_sync(0xffffffffu, o[i], 16);
    }
    if (sub != 0) lsum = 0.f;  // each token's p was added by its 4 lanes; count it once
    lsum = warp_sum(lsum);
    __syncthreads();  // s_red reuse
    if (lane == 0) s_red[warp] = lsum;
    if (lane < 4) {
#pragma unroll
        for (int i = 0; i < EPL; ++i) s_o[warp][lane * EPL + i] = o[i];
    }
    __syncthreads();
    if (tid < D) {
        float acc = 0.f;
#pragma unroll
        for (int w = 0; w < DA_THREADS / 32; ++w) acc += s_o[w][tid];
        wrow[2 + tid] = acc;
    }
    if (tid == 0) {
        float l = 0.f;
#pragma unroll
        for (int w = 0; w < DA_THREADS / 32; ++w) l += s_red[w];
        wrow[0] = bmax;
        wrow[1] = l;
    }
    da_finish<D>(ws, counters, o_out, h, nsplit, tid);
}

template <int D>
__global__ void __launch_bounds__(DA_THREADS)
decode_attn_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ kc,
                   const __nv_bfloat16* __restrict__ vc, float* ws, const int* __restrict__ ctx_len_dev,
                   int heads, int kv_heads, int max_ctx, float scale, int* counters, __nv_bfloat16* o_out) {
    pdl_launch_dependents();
    pdl_wait();
    decode_attn_body<D>(q, kc, vc, ws, ctx_len_dev, heads, kv_heads, max_ctx, scale, counters, o_out, blockIdx.y, gridDim.y);
}

// the same step for up to 4 sequences in ONE launch (gvl_lm_decode_batch): grid (heads, max splits, sequences); every sequence has
// its own cache, context length, workspace and split count (a function of ITS max_ctx)
template <int D>
__global__ void __launch_bounds__(DA_THREADS)
decode_attn_batch_kernel(DecodeAttnBatch b, int heads, int kv_heads, float scale) {
    pdl_launch_dependents();
    pdl_wait();
    const DecodeAttnSeq& p = b.s[blockIdx.z];
    const int nsplit = (p.max_ctx + DA_CHUNK - 1) / DA_CHUNK;
    if ((int)blockIdx.y >= nsplit) return;
    decode_attn_body<D>(p.q, p.kc, p.vc, p.ws, p.ctx_len, heads, kv_heads, p.max_ctx, scale,
                        reinterpret_cast<int*>(p.ws + (size_t)heads * nsplit * (D + 2)), p.o, blockIdx.y, nsplit);
}

// ---------------------------------------------------------------- greedy sampling / step state
// first maximal index (torch.argmax tie order); single CTA.
__global__ void __launch_bounds__(1024)
argmax_kernel(const float* __restrict__ logits, int n, long long* __restrict__ out) {
    __shared__ float s_v[32];
    __shared__ int s_i[32];
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float v = logits[i];
        if (v > best || (v == best && i < bi)) { best = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { s_v[threadIdx.x >> 5] = best; s_i[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x < 32) {
        best = threadIdx.x < (blockDim.x >> 5) ? s_v[threadIdx.x] : -INFINITY;
        bi = threadIdx.x < (blockDim.x >> 5) ? s_i[threadIdx.x] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, best, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (threadIdx.x == 0) out[0] = bi < n ? bi : 0;      // all-NaN logits: a valid id instead of the 0x7fffffff sentinel
    }
}

// x = embed[cur_token]; also bumps nothing. One CTA.
__global__ void embed_token_kernel(const __nv_bfloat16* __restrict__ table, const DecodeState* __restrict__ st,
                                   __nv_bfloat16* __restrict__ x, int dim, int vocab) {
    pdl_launch_dependents();
    pdl_wait();
    long long tok = st->cur_token;
    tok = tok < 0 ? 0 : (tok >= vocab ? vocab - 1 : tok);   // ids handed in through the C ABI are not trusted
    const uint4* src = reinterpret_cast<const uint4*>(table + (size_t)tok * dim);
    for (int i = threadIdx.x; i < dim / 8; i += blockDim.x) reinterpret_cast<uint4*>(x)[i] = src[i];
}

// RoPE for the single new token at position ctx_len (before increment), append k,v to the cache.
__device__ __forceinline__ void rope_decode_body(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ q_out,
                                   __nv_bfloat16* __restrict__ k_cache, __nv_bfloat16* __restrict__ v_cache,
                                   const __nv_bfloat16* __restrict__ cosb, const __nv_bfloat16* __restrict__ sinb,
                                   const DecodeState* __restrict__ st, int heads, int kv_heads, int D, int max_ctx) {
    const int half = D / 2;
    const int total = (heads + 2 * kv_heads) * half;
    const int pos = st->ctx_len;  // slot == position (single unpadded sequence)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int hh = i / half, j = i % half;
        const __nv_bfloat16* src = qkv + (size_t)hh * D;
        const float x1 = __bfloat162float(src[j]), x2 = __bfloat162float(src[j + half]);
        if (hh < heads + kv_heads) {
            const float c1 = __bfloat162float(cosb[(size_t)pos * D + j]);
            const float s1 = __bfloat162float(sinb[(size_t)pos * D + j]);
            const float c2 = __bfloat162float(cosb[(size_t)pos * D + j + half]);
            const float s2 = __bfloat162float(sinb[(size_t)pos * D + j + half]);
            const __nv_bfloat16 y1 = __float2bfloat16_rn(bf16r(x1 * c1) + bf16r(-x2 * s1));
            const __nv_bfloat16 y2 = __float2bfloat16_rn(bf16r(x2 * c2) + bf16r(x1 * s2));
            if (hh < heads) {
                q_out[(size_t)hh * D + j] = y1; q_out[(size_t)hh * D + j + half] = y2;
            } else {
                __nv_bfloat16* dst = k_cache + ((size_t)(hh - heads) * max_ctx + pos) * D;
                dst[j] = y1; dst[j + half] = y2;
            }
        } else {
            __nv_bfloat16* dst = v_cache + ((size_t)(hh - heads - kv_heads) * max_ctx + pos) * D;
            dst[j] = src[j]; dst[j + half] = src[j + half];
        }
    }
}

__global__ void rope_decode_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ q_out,
                                   __nv_bfloat16* __restrict__ k_cache, __nv_bfloat16* __restrict__ v_cache,
                                   const __nv_bfloat16* __restrict__ cosb, const __nv_bfloat16* __restrict__ sinb,
                                   const DecodeState* __restrict__ st, int heads, int kv_heads, int D, int max_ctx) {
    pdl_launch_dependents();
    pdl_wait();
    rope_decode_body(qkv, q_out, k_cache, v_cache, cosb, sinb, st, heads, kv_heads, D, max_ctx);
}
// grid (blocks, sequences): every sequence rotates with ITS position and RoPE table and appends to ITS cache
__global__ void rope_decode_batch_kernel(DecodeRopeBatch b, int heads, int kv_heads, int D) {
    pdl_launch_dependents();
    pdl_wait();
    const DecodeRopeSeq& p = b.s[blockIdx.y];
    rope_decode_body(p.qkv, p.q_out, p.k_cache, p.v_cache, p.cosb, p.sinb, p.st, heads, kv_heads, D, p.max_ctx);
}

// After the per-layer rope kernels of a step: attention must see ctx_len+1 tokens. We keep two counters:
// ctx_len (tokens already in the cache = position of the token being processed) and attn_len = ctx_len+1.
__global__ void step_begin_kernel(DecodeState* st) {
    pdl_launch_dependents();
    pdl_wait();
    st->attn_len = st->ctx_len + 1;
}

// greedy pick + bookkeeping: tokens_out[step] = argmax (or pad after EOS), cur_token = it, ctx_len++, step++.
__global__ void __launch_bounds__(1024)
step_end_kernel(const float* __restrict__ logits, int n, DecodeState* st, long long* __restrict__ tokens_out,
                float* __restrict__ logits_out, long long eos_id, long long pad_id) {
    __shared__ float s_v[32];
    __shared__ int s_i[32];
    pdl_launch_dependents();
    pdl_wait();
    float best = -INFINITY;
    int bi = 0x7fffffff;
    const int step = st->step;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float v = logits[i];
        if (logits_out) logits_out[(size_t)step * n + i] = v;
        if (v > best || (v == best && i < bi)) { best = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { s_v[threadIdx.x >> 5] = best; s_i[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x < 32) {
        best = s_v[threadIdx.x];
        bi = s_i[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, best, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (threadIdx.x == 0) {
            long long tok = bi < n ? bi : 0;                  // all-NaN logits: a valid id
            if (st->finished) tok = pad_id;
            else if (eos_id >= 0 && tok == eos_id) st->finished = 1;
            if (tokens_out) tokens_out[step] = tok;
            st->cur_token = tok;
            st->ctx_len = st->ctx_len + 1;
            st->step = step + 1;
        }
    }
}

}  // namespace



// workspace: [heads][nsplit][D+2] fp32 partials followed by [heads] int32 arrival counters.
// The counters must be ZERO before the first call (gvl_lm_create memsets them; they reset themselves afterwards).
size_t decode_attention_workspace(int heads, int head_dim, int max_ctx) {
    const int nsplit = (max_ctx + DA_CHUNK - 1) / DA_CHUNK;
    return (size_t)heads * nsplit * (head_dim + 2) * sizeof(float) + (size_t)heads * sizeof(int);
}

int decode_attention(const __nv_bfloat16* q, const __nv_bfloat16* kc, const __nv_bfloat16* vc, __nv_bfloat16* o,
                     float* ws, const int* ctx_len_dev, int heads, int kv_heads, int head_dim, int max_ctx,
                     float scale, cudaStream_t s) {
    const int nsplit = (max_ctx + DA_CHUNK - 1) / DA_CHUNK;
    int* counters = reinterpret_cast<int*>(ws + (size_t)heads * nsplit * (head_dim + 2));
    dim3 grid(heads, nsplit);
    cudaError_t e;
    if (head_dim == 64) e = launch_k(decode_attn_kernel<64>, grid, dim3(DA_THREADS), 0, s, q, kc, vc, ws, ctx_len_dev, heads, kv_heads, max_ctx, scale, counters, o);
    else if (head_dim == 96) e = launch_k(decode_attn_kernel<96>, grid, dim3(DA_THREADS), 0, s, q, kc, vc, ws, ctx_len_dev, heads, kv_heads, max_ctx, scale, counters, o);
    else if (head_dim == 128) e = launch_k(decode_attn_kernel<128>, grid, dim3(DA_THREADS), 0, s, q, kc, vc, ws, ctx_len_dev, heads, kv_heads, max_ctx, scale, counters, o);
    else return GVL_ERR_ARG;
    g_launch_count++;
    return (e == cudaSuccess && cudaGetLastError() == cudaSuccess) ? GVL_OK : GVL_ERR_CUDA;
}

int decode_attention_batch(const DecodeAttnBatch& b, int n_seq, int heads, int kv_heads, int head_dim, float scale, cudaStream_t s) {
    if (n_seq < 1 || n_seq > 4) return GVL_ERR_ARG;
    int nsplit = 1;
    for (int i = 0; i < n_seq; ++i) nsplit = max(nsplit, (b.s[i].max_ctx + DA_CHUNK - 1) / DA_CHUNK);
    dim3 grid(heads, nsplit, n_seq);
    cudaError_t e;
    if (head_dim == 64) e = launch_k(decode_attn_batch_kernel<64>, grid, dim3(DA_THREADS), 0, s, b, heads, kv_heads, scale);
    else if (head_dim == 96) e = launch_k(decode_attn_batch_kernel<96>, grid, dim3(DA_THREADS), 0, s, b, heads, kv_heads, scale);
    else if (head_dim == 128) e = launch_k(decode_attn_batch_kernel<128>, grid, dim3(DA_THREADS), 0, s, b, heads, kv_heads, scale);
    else return GVL_ERR_ARG;
    g_launch_count++;
    return (e == cudaSuccess && cudaGetLastError() == cudaSuccess) ? GVL_OK : GVL_ERR_CUDA;
}

int rope_decode_batch(const DecodeRopeBatch& b, int n_seq, int heads, int kv_heads, int D, cudaStream_t s) {
    if (n_seq < 1 || n_seq > 4) return GVL_ERR_ARG;
    const int total = (heads + 2 * kv_heads) * (D / 2);
    launch_k(rope_decode_batch_kernel, dim3((total + 255) / 256, n_seq), dim3(256), 0, s, b, heads, kv_heads, D);
    g_launch_count++;
    return cudaGetLastError() == cudaSuccess ? GVL_OK : GVL_ERR_CUDA;
}

int argmax_f32(const float* logits, int n, long long* out, cudaStream_t s) {
    argmax_kernel<<<1, 1024, 0, s>>>(logits, n, out);
    g_launch_count++;
    return cudaGetLastError() == cudaSuccess ? GVL_OK : GVL_ERR_CUDA;
}

int embed_token(const __nv_bfloat16* table, const DecodeState* st, __nv_bfloat16* x, int dim, int vocab, cudaStream_t s) {
    launch_k(embed_token_kernel, dim3(1), dim3(256), 0, s, table, st, x, dim, vocab);
    g_launch_count++;
    return cudaGetLastError() == cudaSuccess ? GVL_OK : GVL_ERR_CUDA;
}

int rope_decode(const __nv_bfloat16* qkv, __nv_bfloat16* q_out, __nv_bfloat16* k_cache, __nv_bfloat16* v_cache,
                const __nv_bfloat16* cosb, const __nv_bfloat16* sinb, const DecodeState* st, int heads, int kv_heads,
                int D, int max_ctx, cudaStream_t s) {
    const int total = (heads + 2 * kv_heads) * (D / 2);
    launch_k(rope_decode_kernel, dim3((total + 255) / 256), dim3(256), 0, s, qkv, q_out, k_cache, v_cache, cosb, sinb, st, heads, kv_heads, D, max_ctx);
    g_launch_count++;
    return cudaGetLastError() == cudaSuccess ? GVL_OK : GVL_ERR_CUDA;
}

int step_begin(DecodeState* st, cudaStream_t s) {
    launch_k(step_begin_kernel, dim3(1), dim3(1), 0, s, st);
    g_launch_count++;
    return cudaGetLastError() == cudaSuccess ? GVL_OK : GVL_ERR_CUDA;
}

int step_end(const float* logits, int n, DecodeState* st, long long* tokens_out, float* logits_out, long long eos_id,
             long long pad_id, cudaStream_t s) {
    launch_k(step_end_kernel, dim3(1), dim3(1024), 0, s, logits, n, st, tokens_out, logits_out, eos_id, pad_id);
    g_launch_count++;
    return cudaGetLastError() == cudaSuccess ? GVL_OK : GVL_ERR_CUDA;
}

}  // namespace gvl
