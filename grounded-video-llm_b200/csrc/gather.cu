// Index-arithmetic kernels (gather / scatter / pooling). All index maps are bit-exact restatements
// of the reference's reshape/permute/cat chains; the file:line each one follows is given inline.
#include "gvl_internal.h"
#include "ptx.cuh"

namespace gvl {

namespace {

#define GVL_LAUNCH_CHECK() \
    do { g_launch_count++; return cudaGetLastError() == cudaSuccess ? GVL_OK : GVL_ERR_CUDA; } while (0)

// ---- patch-embed feeder: non-overlapping 14x14 conv == GEMM over im2col rows.
// CLIP Conv2d(3,1024,k=s=14) modeling_clip.py:169-175,185 ; IV2 Conv3d(3,1408,(1,14,14)) internvideo2.py:714-722.
// pix: [n_img, C, frames, hw, hw] (frames=1 for CLIP). out: [(n_img*frames*g*g), kpad] bf16,
// k = c*196 + ky*14 + kx (the conv weight's own flattening), zero padded to kpad.
template <typename T>
__global__ void im2col_kernel(const T* __restrict__ pix, __nv_bfloat16* __restrict__ out, int n_img, int chans,
                              int frames, int hw, int kpad) {
    const int g = hw / 14;
    const int kreal = chans * 196;
    const int chunks = kpad / 8;
    const long long total = (long long)n_img * frames * g * g * chunks;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int ch = int(i % chunks);
        long long row = i / chunks;
        const int gx = int(row % g); row /= g;
        const int gy = int(row % g); row /= g;
        const int f = int(row % frames);
        const int n = int(row / frames);
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = ch * 8 + j;
            float x = 0.f;
            if (k < kreal) {
                const int c = k / 196, rem = k % 196, ky = rem / 14, kx = rem % 14;
                const long long src = (((long long)(n * chans + c) * frames + f) * hw + (gy * 14 + ky)) * hw + gx * 14 + kx;
                x = float(pix[src]);
            }
            v[j] = x;
        }
        uint4 o;
        o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]);
        o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]);
        reinterpret_cast<uint4*>(out)[i] = o;
    }
}

// ---- CLIPVisionEmbeddings.forward tail (modeling_clip.py:187-190): cat(class_embedding, patches) + position_embedding,
// promoted to fp32 (cat of fp32 cls with bf16 conv output under autocast).
__global__ void clip_assemble_kernel(const __nv_bfloat16* __restrict__ patch, const float* __restrict__ cls,
                                     const float* __restrict__ pos, float* __restrict__ x, int n_img, int n_patch,
                                     int dim) {
    const int vec = dim / 4;
    const long long total = (long long)n_img * (n_patch + 1) * vec;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c4 = int(i % vec);
        const long long tokg = i / vec;
        const int tok = int(tokg % (n_patch + 1));
        const int n = int(tokg / (n_patch + 1));
        float4 p = __ldg(reinterpret_cast<const float4*>(pos + (size_t)tok * dim) + c4);
        float4 v;
        if (tok == 0) {
            v = __ldg(reinterpret_cast<const float4*>(cls) + c4);
        } else {
            uint2 raw = *(reinterpret_cast<const uint2*>(patch + ((size_t)n * n_patch + tok - 1) * dim) + c4);
            float2 a = unpack_bf16(raw.x), b = unpack_bf16(raw.y);
            v = make_float4(a.x, a.y, b.x, b.y);
        }
        reinterpret_cast<float4*>(x)[i] = make_float4(v.x + p.x, v.y + p.y, v.z + p.z, v.w + p.w);
    }
}

// ---- PretrainInternVideo2.forward prologue (internvideo2.py:975-1005): cat(cls_token, patches) + pos_embed, bf16.
__global__ void iv2_assemble_kernel(const __nv_bfloat16* __restrict__ patch, const __nv_bfloat16* __restrict__ cls,
                                    const __nv_bfloat16* __restrict__ pos, __nv_bfloat16* __restrict__ x, int n_seg,
                                    int n_patch, int dim) {
    const int vec = dim / 8;
    const long long total = (long long)n_seg * (n_patch + 1) * vec;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c8 = int(i % vec);
        const long long tokg = i / vec;
        const int tok = int(tokg % (n_patch + 1));
        const int n = int(tokg / (n_patch + 1));
        uint4 p = __ldg(reinterpret_cast<const uint4*>(pos + (size_t)tok * dim) + c8);
        uint4 v = (tok == 0) ? __ldg(reinterpret_cast<const uint4*>(cls) + c8)
                             : *(reinterpret_cast<const uint4*>(patch + ((size_t)n * n_patch + tok - 1) * dim) + c8);
        uint4 o;
        float2 a, b;
        a = unpack_bf16(v.x); b = unpack_bf16(p.x); o.x = pack_bf16(a.x + b.x, a.y + b.y);
        a = unpack_bf16(v.y); b = unpack_bf16(p.y); o.y = pack_bf16(a.x + b.x, a.y + b.y);
        a = unpack_bf16(v.z); b = unpack_bf16(p.z); o.z = pack_bf16(a.x + b.x, a.y + b.y);
        a = unpack_bf16(v.w); b = unpack_bf16(p.w); o.w = pack_bf16(a.x + b.x, a.y + b.y);
        reinterpret_cast<uint4*>(x)[i] = o;
    }
}

// ---- reshape_hd_patches_2x2merge_phi3 + add_image_newline_phi3 (llava_next_video.py:454-489), h_crop=w_crop=1.
// hs: CLIP hidden_states[-2], fp32 [n_img, 577, 1024] (row 0 = cls, dropped at :505).
// out: bf16 [n_img, 12*13, 4096]; out[n, y*13+x, (dy*2+dx)*1024+c] = hs[n, 1+(2y+dy)*24+(2x+dx), c]; x==12 -> sub_GN.
__global__ void hd_merge_kernel(const float* __restrict__ hs, const float* __restrict__ sub_gn,
                                __nv_bfloat16* __restrict__ out, int n_img) {
    const long long total = (long long)n_img * 156 * 512;  // 8-element chunks
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int ch = int(i % 512);
        const long long tokg = i / 512;
        const int tok = int(tokg % 156);
        const int n = int(tokg / 156);
        const int y = tok / 13, x = tok % 13;
        const int col = ch * 8;
        const float* src;
        if (x == 12) {
            src = sub_gn + col;
        } else {
            const int quad = col >> 10, c = col & 1023;
            const int dy = quad >> 1, dx = quad & 1;
            src = hs + ((size_t)n * 577 + 1 + (2 * y + dy) * 24 + (2 * x + dx)) * 1024 + c;
        }
        float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
        uint4 o;
        o.x = pack_bf16(a.x, a.y); o.y = pack_bf16(a.z, a.w);
        o.z = pack_bf16(b.x, b.y); o.w = pack_bf16(b.z, b.w);
        reinterpret_cast<uint4*>(out)[i] = o;
    }
}

// ---- AdaptiveAvgPool3d([T,4,4]) on the 16x16 per-frame grid (llava_next_video.py:544-549): exact 4x4 means.
// x: bf16 [n_seg, 1+frames*256, dim] (cls dropped at :532); out: bf16 [n_seg, frames*16, dim], token (f, py, px).
__global__ void iv2_pool_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, int n_seg,
                                int frames, int dim) {
    const int vec = dim / 8;
    const long long total = (long long)n_seg * frames * 16 * vec;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c8 = int(i % vec);
        long long t = i / vec;
        const int cell = int(t % 16); t /= 16;
        const int f = int(t % frames);
        const int n = int(t / frames);
        const int py = cell / 4, px = cell % 4;
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const __nv_bfloat16* base = x + ((size_t)n * (1 + frames * 256) + 1 + f * 256) * dim;
#pragma unroll
        for (int dy = 0; dy < 4; ++dy)
#pragma unroll
            for (int dx = 0; dx < 4; ++dx) {
                uint4 v = *(reinterpret_cast<const uint4*>(base + (size_t)((4 * py + dy) * 16 + 4 * px + dx) * dim) + c8);
                float2 a;
                a = unpack_bf16(v.x); acc[0] += a.x; acc[1] += a.y;
                a = unpack_bf16(v.y); acc[2] += a.x; acc[3] += a.y;
                a = unpack_bf16(v.z); acc[4] += a.x; acc[5] += a.y;
                a = unpack_bf16(v.w); acc[6] += a.x; acc[7] += a.y;
            }
        uint4 o;
        o.x = pack_bf16(acc[0] * 0.0625f, acc[1] * 0.0625f); o.y = pack_bf16(acc[2] * 0.0625f, acc[3] * 0.0625f);
        o.z = pack_bf16(acc[4] * 0.0625f, acc[5] * 0.0625f); o.w = pack_bf16(acc[6] * 0.0625f, acc[7] * 0.0625f);
        reinterpret_cast<uint4*>(out)[i] = o;
    }
}

// ---- Llama path: AdaptiveAvgPool3d([segs,8,8]) on the 24x24 CLIP grid (llava_next_video.py:509-517): exact 3x3 means.
__global__ void clip_pool3_kernel(const float* __restrict__ hs, __nv_bfloat16* __restrict__ out, int n_img) {
    const long long total = (long long)n_img * 64 * 128;  // 8-element chunks of 1024
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c8 = int(i % 128);
        long long t = i / 128;
        const int cell = int(t % 64);
        const int n = int(t / 64);
        const int py = cell / 8, px = cell % 8;
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int dy = 0; dy < 3; ++dy)
            for (int dx = 0; dx < 3; ++dx) {
                const float* src = hs + ((size_t)n * 577 + 1 + (3 * py + dy) * 24 + 3 * px + dx) * 1024 + c8 * 8;
                float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
                acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
                acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
            }
        const float r = 1.0f / 9.0f;
        uint4 o;
        o.x = pack_bf16(acc[0] * r, acc[1] * r); o.y = pack_bf16(acc[2] * r, acc[3] * r);
        o.z = pack_bf16(acc[4] * r, acc[5] * r); o.w = pack_bf16(acc[6] * r, acc[7] * r);
        reinterpret_cast<uint4*>(out)[i] = o;
    }
}

// ---- per-segment stream concat (llava_next_video.py:563-564): [spatial a_rows | temporal b_rows | newline 1].
__global__ void visual_concat_kernel(const __nv_bfloat16* __restrict__ a, int a_rows, const __nv_bfloat16* __restrict__ b,
                                     int b_rows, const __nv_bfloat16* __restrict__ newline,
                                     __nv_bfloat16* __restrict__ out, int n_seg, int dim) {
    const int vec = dim / 8;
    const int per = a_rows + b_rows + 1;
    const long long total = (long long)n_seg * per * vec;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c8 = int(i % vec);
        const long long rg = i / vec;
        const int r = int(rg % per);
        const int n = int(rg / per);
        const uint4* src;
        if (r < a_rows) src = reinterpret_cast<const uint4*>(a + ((size_t)n * a_rows + r) * dim);
        else if (r < a_rows + b_rows) src = reinterpret_cast<const uint4*>(b + ((size_t)n * b_rows + (r - a_rows)) * dim);
        else src = reinterpret_cast<const uint4*>(newline);
        reinterpret_cast<uint4*>(out)[i] = src[c8];
    }
}

// ---- prepare_multimodal_inputs (llava_next_video.py:568-596): embed_tokens(ids[:p]) ++ visual ++ embed_tokens(ids[p+1:])
// ('text' samples: visual appended last). ids has t_text entries with the -200 sentinel at img_pos.
__global__ void embed_splice_kernel(const long long* __restrict__ ids, int t_text, int img_pos,
                                    const __nv_bfloat16* __restrict__ table, const __nv_bfloat16* __restrict__ visual,
                                    int n_vis, __nv_bfloat16* __restrict__ out, int dim, int vis_last) {
    const int vec = dim / 8;
    const int S = t_text - 1 + n_vis;
    const long long total = (long long)S * vec;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c8 = int(i % vec);
        const int s = int(i / vec);
        const uint4* src;
        if (!vis_last) {
            if (s < img_pos) src = reinterpret_cast<const uint4*>(table + (size_t)ids[s] * dim);
            else if (s < img_pos + n_vis) src = reinterpret_cast<const uint4*>(visual + (size_t)(s - img_pos) * dim);
            else src = reinterpret_cast<const uint4*>(table + (size_t)ids[s - n_vis + 1] * dim);
        } else {
            if (s < img_pos) src = reinterpret_cast<const uint4*>(table + (size_t)ids[s] * dim);
            else if (s < t_text - 1) src = reinterpret_cast<const uint4*>(table + (size_t)ids[s + 1] * dim);
            else src = reinterpret_cast<const uint4*>(visual + (size_t)(s - (t_text - 1)) * dim);
        }
        reinterpret_cast<uint4*>(out)[i] = src[c8];
    }
}

// ---- RoPE on q,k + KV-cache append (modeling_phi3.py:413-445 apply_rotary_pos_emb, :721 cache.update;
// modeling_llama.py:173-204, :451). cos/sin tables are bf16 [max_pos, D] (fp32 trig, cast to bf16: phi3 :409).
// q_embed = bf16(bf16(q*cos) + bf16(rot(q)*sin)), rot = cat(-x2, x1) with half-split pairing.
// qkv: [T, (H+2*KVH)*D]; q_out: [T, H*D]; caches: [KVH, max_ctx, D] written at slot positions[t] (or pos0+t).
__global__ void rope_cache_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ q_out,
                                  __nv_bfloat16* __restrict__ k_cache, __nv_bfloat16* __restrict__ v_cache,
                                  const __nv_bfloat16* __restrict__ cosb, const __nv_bfloat16* __restrict__ sinb,
                                  const int* __restrict__ positions, int tokens, int heads, int kv_heads, int D,
                                  int pos0, int max_ctx) {
    const int half = D / 2;
    const int per_tok = (heads + 2 * kv_heads) * half;  // pair-threads for q,k; v handled as pairs too
    const long long total = (long long)tokens * per_tok;
    const int ld = (heads + 2 * kv_heads) * D;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int t = int(i / per_tok);
        const int r = int(i % per_tok);
        const int hh = r / half, j = r % half;
        const int slot = pos0 + t;
        const int pos = positions ? positions[t] : slot;
        const __nv_bfloat16* src = qkv + (size_t)t * ld + (size_t)hh * D;
        const float x1 = __bfloat162float(src[j]), x2 = __bfloat162float(src[j + half]);
        if (hh < heads + kv_heads) {
            const float c1 = __bfloat162float(cosb[(size_t)pos * D + j]);
            const float s1 = __bfloat162float(sinb[(size_t)pos * D + j]);
            const float c2 = __bfloat162float(cosb[(size_t)pos * D + j + half]);
            const float s2 = __bfloat162float(sinb[(size_t)pos * D + j + half]);
            const __nv_bfloat16 y1 = __float2bfloat16_rn(bf16r(x1 * c1) + bf16r(-x2 * s1));
            const __nv_bfloat16 y2 = __float2bfloat16_rn(bf16r(x2 * c2) + bf16r(x1 * s2));
            if (hh < heads) {
                __nv_bfloat16* dst = q_out + (size_t)t * heads * D + (size_t)hh * D;
                dst[j] = y1; dst[j + half] = y2;
            } else {
                __nv_bfloat16* dst = k_cache + ((size_t)(hh - heads) * max_ctx + slot) * D;
                dst[j] = y1; dst[j + half] = y2;
            }
        } else {
            __nv_bfloat16* dst = v_cache + ((size_t)(hh - heads - kv_heads) * max_ctx + slot) * D;
            dst[j] = src[j]; dst[j + half] = src[j + half];
        }
    }
}

// 128-bit vectorised variant (head_dim % 16 == 0): one thread owns 8 consecutive elements j..j+7 of the first half
// of a head and their partners j+D/2.. of the second half; same arithmetic and rounding points as above.
__device__ __forceinline__ void rope8(const uint4& a, const uint4& b, const uint4& c1, const uint4& s1, const uint4& c2,
                                      const uint4& s2, uint4& y1, uint4& y2) {
    const uint32_t* pa = reinterpret_cast<const uint32_t*>(&a);
    const uint32_t* pb = reinterpret_cast<const uint32_t*>(&b);
    const uint32_t* pc1 = reinterpret_cast<const uint32_t*>(&c1);
    const uint32_t* ps1 = reinterpret_cast<const uint32_t*>(&s1);
    const uint32_t* pc2 = reinterpret_cast<const uint32_t*>(&c2);
    const uint32_t* ps2 = reinterpret_cast<const uint32_t*>(&s2);
    uint32_t* o1 = reinterpret_cast<uint32_t*>(&y1);
    uint32_t* o2 = reinterpret_cast<uint32_t*>(&y2);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 x1 = unpack_bf16(pa[i]), x2 = unpack_bf16(pb[i]);
        const float2 cc1 = unpack_bf16(pc1[i]), ss1 = unpack_bf16(ps1[i]);
        const float2 cc2 = unpack_bf16(pc2[i]), ss2 = unpack_bf16(ps2[i]);
        o1[i] = pack_bf16(bf16r(x1.x * cc1.x) + bf16r(-x2.x * ss1.x), bf16r(x1.y * cc1.y) + bf16r(-x2.y * ss1.y));
        o2[i] = pack_bf16(bf16r(x2.x * cc2.x) + bf16r(x1.x * ss2.x), bf16r(x2.y * cc2.y) + bf16r(x1.y * ss2.y));
    }
}

__global__ void rope_cache_vec_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ q_out,
                                      __nv_bfloat16* __restrict__ k_cache, __nv_bfloat16* __restrict__ v_cache,
                                      const __nv_bfloat16* __restrict__ cosb, const __nv_bfloat16* __restrict__ sinb,
                                      const int* __restrict__ positions, int tokens, int heads, int kv_heads, int D,
                                      int pos0, int max_ctx) {
    const int half = D / 2;
    const int cph = half / 8;                          // 8-element chunks per half head
    const int per_tok = (heads + 2 * kv_heads) * cph;
    const long long total = (long long)tokens * per_tok;
    const int ld = (heads + 2 * kv_heads) * D;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int t = int(i / per_tok);
        const int r = int(i % per_tok);
        const int hh = r / cph, j = (r % cph) * 8;
        const int slot = pos0 + t;
        const int pos = positions ? positions[t] : slot;
        const __nv_bfloat16* src = qkv + (size_t)t * ld + (size_t)hh * D;
        const uint4 a = *reinterpret_cast<const uint4*>(src + j);
        const uint4 b = *reinterpret_cast<const uint4*>(src + j + half);
        if (hh < heads + kv_heads) {
            const __nv_bfloat16* cp = cosb + (size_t)pos * D;
            const __nv_bfloat16* sp = sinb + (size_t)pos * D;
            uint4 y1, y2;
            rope8(a, b, __ldg(reinterpret_cast<const uint4*>(cp + j)), __ldg(reinterpret_cast<const uint4*>(sp + j)),
                  __ldg(reinterpret_cast<const uint4*>(cp + j + half)), __ldg(reinterpret_cast<const uint4*>(sp + j + half)),
                  y1, y2);
            __nv_bfloat16* dst = (hh < heads) ? q_out + (size_t)t * heads * D + (size_t)hh * D
                                              : k_cache + ((size_t)(hh - heads) * max_ctx + slot) * D;
            *reinterpret_cast<uint4*>(dst + j) = y1;
            *reinterpret_cast<uint4*>(dst + j + half) = y2;
        } else {
            __nv_bfloat16* dst = v_cache + ((size_t)(hh - heads - kv_heads) * max_ctx + slot) * D;
            *reinterpret_cast<uint4*>(dst + j) = a;
            *reinterpret_cast<uint4*>(dst + j + half) = b;
        }
    }
}

inline int grid_for(long long total, int block) {
    long long g = (total + block - 1) / block;
    long long cap = (long long)num_sms() * 16;
    return int(g < cap ? (g < 1 ? 1 : g) : cap);
}

}  // namespace

int im2col_patch14(const void* pix, int pix_is_f32, __nv_bfloat16* out, int n_img, int chans, int frames, int hw,
                   int kpad, cudaStream_t s) {
    if (hw % 14 != 0 || kpad % 8 != 0 || kpad < chans * 196) return GVL_ERR_ARG;
    const int g = hw / 14;
    const long long total = (long long)n_img * frames * g * g * (kpad / 8);
    if (pix_is_f32)
        im2col_kernel<float><<<grid_for(total, 256), 256, 0, s>>>(reinterpret_cast<const float*>(pix), out, n_img, chans, frames, hw, kpad);
    else
        im2col_kernel<__nv_bfloat16><<<grid_for(total, 256), 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(pix), out, n_img, chans, frames, hw, kpad);
    GVL_LAUNCH_CHECK();
}

int clip_assemble(const __nv_bfloat16* patch, const float* cls, const float* pos, float* x, int n_img, int n_patch,
                  int dim, cudaStream_t s) {
    if (dim % 4 != 0) return GVL_ERR_ARG;
    const long long total = (long long)n_img * (n_patch + 1) * (dim / 4);
    clip_assemble_kernel<<<grid_for(total, 256), 256, 0, s>>>(patch, cls, pos, x, n_img, n_patch, dim);
    GVL_LAUNCH_CHECK();
}

int iv2_assemble(const __nv_bfloat16* patch, const __nv_bfloat16* cls, const __nv_bfloat16* pos, __nv_bfloat16* x,
                 int n_seg, int n_patch, int dim, cudaStream_t s) {
    if (dim % 8 != 0) return GVL_ERR_ARG;
    const long long total = (long long)n_seg * (n_patch + 1) * (dim / 8);
    iv2_assemble_kernel<<<grid_for(total, 256), 256, 0, s>>>(patch, cls, pos, x, n_seg, n_patch, dim);
    GVL_LAUNCH_CHECK();
}

int hd_merge_newline(const float* hs, const float* sub_gn, __nv_bfloat16* out, int n_img, cudaStream_t s) {
    const long long total = (long long)n_img * 156 * 512;
    hd_merge_kernel<<<grid_for(total, 256), 256, 0, s>>>(hs, sub_gn, out, n_img);
    GVL_LAUNCH_CHECK();
}

int iv2_pool(const __nv_bfloat16* x, __nv_bfloat16* out, int n_seg, int frames, int dim, cudaStream_t s) {
    if (dim % 8 != 0) return GVL_ERR_ARG;
    const long long total = (long long)n_seg * frames * 16 * (dim / 8);
    iv2_pool_kernel<<<grid_for(total, 256), 256, 0, s>>>(x, out, n_seg, frames, dim);
    GVL_LAUNCH_CHECK();
}

int clip_pool3(const float* hs, __nv_bfloat16* out, int n_img, cudaStream_t s) {
    const long long total = (long long)n_img * 64 * 128;
    clip_pool3_kernel<<<grid_for(total, 256), 256, 0, s>>>(hs, out, n_img);
    GVL_LAUNCH_CHECK();
}

int visual_concat(const __nv_bfloat16* a, int a_rows, const __nv_bfloat16* b, int b_rows,
                  const __nv_bfloat16* newline, __nv_bfloat16* out, int n_seg, int dim, cudaStream_t s) {
    if (dim % 8 != 0) return GVL_ERR_ARG;
    const long long total = (long long)n_seg * (a_rows + b_rows + 1) * (dim / 8);
    visual_concat_kernel<<<grid_for(total, 256), 256, 0, s>>>(a, a_rows, b, b_rows, newline, out, n_seg, dim);
    GVL_LAUNCH_CHECK();
}

int embed_splice(const long long* ids, int t_text, int img_pos, const __nv_bfloat16* table,
                 const __nv_bfloat16* visual, int n_vis, __nv_bfloat16* out, int dim, int vis_last, cudaStream_t s) {
    if (dim % 8 != 0 || img_pos < 0 || img_pos >= t_text) return GVL_ERR_ARG;
    const long long total = (long long)(t_text - 1 + n_vis) * (dim / 8);
    embed_splice_kernel<<<grid_for(total, 256), 256, 0, s>>>(ids, t_text, img_pos, table, visual, n_vis, out, dim, vis_last);
    GVL_LAUNCH_CHECK();
}

int rope_qkv_cache(const __nv_bfloat16* qkv, __nv_bfloat16* q_out, __nv_bfloat16* k_cache, __nv_bfloat16* v_cache,
                   const __nv_bfloat16* cos, const __nv_bfloat16* sin, const int* positions, int tokens, int heads,
                   int kv_heads, int head_dim, int pos0, int max_ctx, cudaStream_t s) {
    if (head_dim % 2 != 0 || pos0 + tokens > max_ctx) return GVL_ERR_ARG;
    if (head_dim % 16 == 0) {
        const long long totv = (long long)tokens * (heads + 2 * kv_heads) * (head_dim / 16);
        rope_cache_vec_kernel<<<grid_for(totv, 256), 256, 0, s>>>(qkv, q_out, k_cache, v_cache, cos, sin, positions,
                                                                  tokens, heads, kv_heads, head_dim, pos0, max_ctx);
        GVL_LAUNCH_CHECK();
    }
    const long long total = (long long)tokens * (heads + 2 * kv_heads) * (head_dim / 2);
    rope_cache_kernel<<<grid_for(total, 256), 256, 0, s>>>(qkv, q_out, k_cache, v_cache, cos, sin, positions, tokens,
                                                          heads, kv_heads, head_dim, pos0, max_ctx);
    GVL_LAUNCH_CHECK();
}

}  // namespace gvl
