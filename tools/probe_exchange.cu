// Micro-benchmark (decode_mega.cu design input, round 2): what does ONE all-to-all vector exchange between 148 persistent CTAs
// cost, as a function of the hand-off primitive, while a producer warp per CTA keeps an HBM bulk-copy stream running?
// Every iteration each CTA writes its interleaved 16-byte chunks of a V-element bf16 vector (chunk j belongs to CTA j % G, like
// the GEMV units of the decode step) and then needs the WHOLE vector in shared memory.
//   mode 0  grid barrier (threadfence + atomicAdd + one polling thread) + ld.global.cg of the vector          [round-1 design]
//   mode 1  same, arrive = red.release.gpu, poll = ld.acquire by a whole warp
//   mode 2  flags in data, 8-byte packets  {2 x bf16, iteration}            (every thread polls the packets it needs)
//   mode 3  flags in data, 16-byte packets {6 x bf16, iteration}
//   mode 4  per-CTA release flags: data stores, fence, st.release flag[cta]; readers poll the 148 flags, CTA barrier, load data
//   mode 5  per-CTA release flags, no CTA barrier: each thread polls the flag of the CTA that owns the chunk it loads
// Build / run:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe_exchange tools/probe_exchange.cu && tools/probe_exchange
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void red_release(unsigned* p, unsigned v) { asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint4 ld_cg4(const void* p) { uint4 r; asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory"); return r; }
__device__ __forceinline__ void st_cg4(void* p, uint4 v) { asm volatile("st.global.cg.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
// 8-byte {data, flag} packets as SCALAR 64-bit relaxed accesses: single-copy atomic by the PTX memory model (a .v2.u32 / .v4.u32
// weak access may be split -- the first version of this probe saw {new flag, old data} with ld.global.cg.v4 / st.global.cg.v2)
__device__ __forceinline__ void st_pkt(void* p, uint32_t data, uint32_t flag) {
    asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(p), "l"((unsigned long long)data | ((unsigned long long)flag << 32)) : "memory");
}
__device__ __forceinline__ unsigned long long ld_pkt(const void* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void cbar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

constexpr int SLOT = 8192, NLANE = 8, MAXSLOTS = 3;

struct Args {
    const uint8_t* stream;      // HBM source of the background stream
    size_t stream_bytes;
    int inflight;               // bulk copies outstanding per producer lane (0 = no background stream)
    int n_iter, V, mode;
    unsigned* counter;          // grid barrier counter
    unsigned* flags;            // [G] per-CTA flags
    uint32_t* vec;              // exchange buffer (two copies, alternating per iteration, large enough for any packet format)
    unsigned long long* out;    // [0] wall ns, [1] errors, [2] streamed bytes
    int work_ns;                // simulated work between exchanges
};

// value of element n in iteration it (bf16 bit pattern; never 0 in the flag position because flags are it + 1)
__device__ __forceinline__ uint32_t val16(int it, int n) { return (uint32_t)((it * 131 + n * 7 + 1) & 0x7fff); }

__global__ void __launch_bounds__(320, 1) exchange_kernel(Args a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full[NLANE * MAXSLOTS];
    __shared__ volatile int s_done;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = gridDim.x, c = blockIdx.x;
    if (tid == 0) {
        for (int s = 0; s < NLANE * MAXSLOTS; ++s) mbar_init(smem_u32(&full[s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_done = 0;
    }
    __syncthreads();
    cooperative_groups::this_grid().sync();                 // every thread takes part, before the roles split
    uint8_t* ring = smem;                                   // [NLANE][MAXSLOTS][SLOT]
    uint16_t* sx = reinterpret_cast<uint16_t*>(smem + NLANE * MAXSLOTS * SLOT);

    if (warp == 8) {
        // ---------------- background HBM stream: lane l keeps `inflight` 8 KB copies outstanding, re-issuing as they land
        if (lane >= NLANE || a.inflight <= 0) return;
        const size_t n_items = a.stream_bytes / SLOT;
        size_t item = ((size_t)c * NLANE + lane) * 4099 % n_items;
        uint32_t issued = 0, landed = 0;
        unsigned long long bytes = 0;
        while (!s_done) {
            while (issued - landed < (uint32_t)a.inflight) {
                const uint32_t slot = issued % MAXSLOTS;
                const uint32_t bar = smem_u32(&full[lane * MAXSLOTS + slot]);
                mbar_expect_tx(bar, SLOT);
                bulk_g2s(smem_u32(ring) + (lane * MAXSLOTS + slot) * SLOT, a.stream + item * SLOT, SLOT, bar);
                item += (size_t)G * NLANE;
                if (item >= n_items) item -= n_items;
                ++issued;
            }
            const uint32_t slot = landed % MAXSLOTS;
            if (mbar_test(smem_u32(&full[lane * MAXSLOTS + slot]), (landed / MAXSLOTS) & 1)) { ++landed; bytes += SLOT; }
        }
        while (landed < issued) {                           // drain
            const uint32_t slot = landed % MAXSLOTS;
            if (mbar_test(smem_u32(&full[lane * MAXSLOTS + slot]), (landed / MAXSLOTS) & 1)) ++landed;
        }
        atomicAdd(a.out + 2, bytes);
        return;
    }
    if (warp > 8) return;

    // ---------------- the 256 exchanging threads
    const int V = a.V, nchunk = V / 8;                      // 16-byte chunks of 8 elements; chunk j belongs to CTA j % G
    const int my_chunks = nchunk > c ? (nchunk - c + G - 1) / G : 0;
    unsigned errors = 0;
    unsigned bars = 0;
    const size_t buf_words = 65536;                         // one exchange buffer copy (uint32 words)
    const unsigned long long t0 = gtime();
    for (int it = 0; it < a.n_iter; ++it) {
        uint32_t* buf = a.vec + (size_t)(it & 1) * buf_words;
        const unsigned flag = (unsigned)it + 1u;
        if (a.work_ns > 0) { const long long w0 = clock64(); while (clock64() - w0 < (long long)a.work_ns * 17 / 10) {} }   // ~1.7 GHz
        // ---- write my chunks
        if (a.mode <= 1 || a.mode == 4 || a.mode == 5) {
            for (int k = tid; k < my_chunks * 4; k += 256) {                  // 4 threads x 4 bytes per chunk (scattered like the epilogue)
                const int j = c + (k >> 2) * G, q = k & 3;
                const int n = j * 8 + q * 2;
                buf[j * 4 + q] = val16(it, n) | (val16(it, n + 1) << 16);
            }
        } else if (a.mode == 2) {
            for (int k = tid; k < my_chunks * 4; k += 256) {                  // one 8-byte packet per 2 elements
                const int j = c + (k >> 2) * G, q = k & 3;
                const int n = j * 8 + q * 2;
                asm volatile("st.global.cg.v2.u32 [%0], {%1,%2};" ::"l"(buf + (size_t)(n >> 1) * 2), "r"(val16(it, n) | (val16(it, n + 1) << 16)), "r"(flag) : "memory");
            }
        } else if (a.mode == 6) {
            for (int k = tid; k < my_chunks * 4; k += 256) {
                const int j = c + (k >> 2) * G, q = k & 3;
                const int n = j * 8 + q * 2;
                st_pkt(buf + (size_t)(n >> 1) * 2, val16(it, n) | (val16(it, n + 1) << 16), flag);
            }
        } else {                                                              // mode 3: 16-byte packets of 6 elements: written by the CTAs that own
            // packet p covers elements 6p .. 6p+5; to keep ownership simple here packet p belongs to CTA p % G
            const int npk = (V + 5) / 6;
            const int mine = npk > c ? (npk - c + G - 1) / G : 0;
            for (int k = tid; k < mine; k += 256) {
                const int p = c + k * G, n = p * 6;
                uint4 v;
                v.x = val16(it, n) | (val16(it, n + 1) << 16); v.y = val16(it, n + 2) | (val16(it, n + 3) << 16);
                v.z = val16(it, n + 4) | (val16(it, n + 5) << 16); v.w = flag;
                st_cg4(buf + (size_t)p * 4, v);
            }
        }
        // ---- hand-off + gather the whole vector into shared memory
        if (a.mode == 0) {
            cbar();
            ++bars;
            if (tid == 0) {
                __threadfence();
                atomicAdd(a.counter, 1u);
                while (ld_acquire(a.counter) < bars * (unsigned)G) {}
            }
            cbar();
            for (int j = tid; j < nchunk; j += 256) reinterpret_cast<uint4*>(sx)[j] = ld_cg4(buf + (size_t)j * 4);
            cbar();
        } else if (a.mode == 1) {
            cbar();
            ++bars;
            if (warp == 0) {
                if (lane == 0) red_release(a.counter, 1u);
                while (ld_acquire(a.counter) < bars * (unsigned)G) {}
            }
            cbar();
            for (int j = tid; j < nchunk; j += 256) reinterpret_cast<uint4*>(sx)[j] = ld_cg4(buf + (size_t)j * 4);
            cbar();
        } else if (a.mode == 2) {
            // thread loads 2 packets (4 elements) per 16-byte load; all loads first, then re-poll the late ones
            const int nld = V / 4;
            for (int q0 = 0; q0 < nld; q0 += 256 * 4) {
                uint4 r[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) { const int q = q0 + u * 256 + tid; if (q < nld) r[u] = ld_cg4(buf + (size_t)q * 4); }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int q = q0 + u * 256 + tid;
                    if (q < nld) {
                        while (r[u].y != flag || r[u].w != flag) r[u] = ld_cg4(buf + (size_t)q * 4);
                        reinterpret_cast<uint2*>(sx)[q] = make_uint2(r[u].x, r[u].z);
                    }
                }
            }
            cbar();
        } else if (a.mode == 3) {
            const int npk = (V + 5) / 6;
            for (int q0 = 0; q0 < npk; q0 += 256 * 4) {
                uint4 r[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) { const int q = q0 + u * 256 + tid; if (q < npk) r[u] = ld_cg4(buf + (size_t)q * 4); }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int q = q0 + u * 256 + tid;
                    if (q < npk) {
                        while (r[u].w != flag) r[u] = ld_cg4(buf + (size_t)q * 4);
                        uint32_t* d = reinterpret_cast<uint32_t*>(sx) + q * 3;
                        d[0] = r[u].x; d[1] = r[u].y; d[2] = r[u].z;
                    }
                }
            }
            cbar();
        } else if (a.mode == 6) {
            const int npk = V / 2;                                            // one packet = 2 elements; all loads first, then re-poll
            for (int q0 = 0; q0 < npk; q0 += 256 * 8) {
                unsigned long long r[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) { const int q = q0 + u * 256 + tid; if (q < npk) r[u] = ld_pkt(buf + (size_t)q * 2); }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int q = q0 + u * 256 + tid;
                    if (q < npk) {
                        while ((unsigned)(r[u] >> 32) != flag) r[u] = ld_pkt(buf + (size_t)q * 2);
                        reinterpret_cast<uint32_t*>(sx)[q] = (uint32_t)r[u];
                    }
                }
            }
            cbar();
        } else if (a.mode == 4) {
            cbar();
            if (tid == 0) { __threadfence(); st_release(a.flags + c, flag); }
            if (tid < G) { while (ld_acquire(a.flags + tid) < flag) {} }
            cbar();
            for (int j = tid; j < nchunk; j += 256) reinterpret_cast<uint4*>(sx)[j] = ld_cg4(buf + (size_t)j * 4);
            cbar();
        } else {
            cbar();
            if (tid == 0) { __threadfence(); st_release(a.flags + c, flag); }
            for (int j = tid; j < nchunk; j += 256) {
                while (ld_acquire(a.flags + (j % G)) < flag) {}
                reinterpret_cast<uint4*>(sx)[j] = ld_cg4(buf + (size_t)j * 4);
            }
            cbar();
        }
        // ---- verify a sample (the whole vector every 16th iteration)
        if ((it & 15) == 0) {
            for (int n = tid; n < V; n += 256) {
                const bool bad = sx[n] != (uint16_t)val16(it, n);
                errors += bad;
                if (bad && atomicAdd(a.out + 3, 1ull) < 6ull)
                    printf("  mismatch mode %d cta %d it %d n %d got %04x want %04x (it-1 %04x it-2 %04x it+1 %04x)\n", a.mode, c, it, n, (unsigned)sx[n],
                           (unsigned)val16(it, n), (unsigned)val16(it - 1, n), (unsigned)val16(it - 2, n), (unsigned)val16(it + 1, n));
            }
        } else {
            const int n = (tid * 37 + it) % V;
            errors += sx[n] != (uint16_t)val16(it, n);
        }
        cbar();                                             // sx is overwritten by the next iteration
    }
    const unsigned long long t1 = gtime();
    if (errors) atomicAdd(a.out + 1, (unsigned long long)errors);
    cbar();
    if (tid == 0) {
        s_done = 1;
        if (c == 0) a.out[0] = t1 - t0;
    }
}

int main(int argc, char** argv) {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int G = prop.multiProcessorCount;
    const size_t stream_bytes = (size_t)2 << 30;
    uint8_t* stream;
    CK(cudaMalloc(&stream, stream_bytes));
    CK(cudaMemset(stream, 1, stream_bytes));
    unsigned *counter, *flags;
    uint32_t* vec;
    unsigned long long* out;
    CK(cudaMalloc(&counter, 4));
    CK(cudaMalloc(&flags, 4 * 1024));
    CK(cudaMalloc(&vec, 2 * 65536 * 4));
    CK(cudaMalloc(&out, 64));
    const size_t smem = NLANE * MAXSLOTS * SLOT + 16384 + 1024;
    CK(cudaFuncSetAttribute(exchange_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int n_iter = argc > 1 ? atoi(argv[1]) : 2000;
    printf("G=%d  n_iter=%d  (ns per exchange = wall / n_iter, globaltimer on CTA 0)\n", G, n_iter);
    const char* names[] = {"barrier+load", "red.release/warp-poll+load", "LL 8B packets", "LL 16B packets", "per-CTA flags+bar+load", "per-CTA flags, per-thread poll", "LL 8B packets, scalar b64 relaxed"};
    for (int V : {3072, 8192}) {
        for (int inflight : {0, 1, 2}) {
            for (int work_ns : {0, 4000}) {
                for (int mode : {0, 2, 6}) {
                    CK(cudaMemset(counter, 0, 4));
                    CK(cudaMemset(flags, 0, 4 * 1024));
                    CK(cudaMemset(vec, 0, 2 * 65536 * 4));
                    CK(cudaMemset(out, 0, 64));
                    Args a{stream, stream_bytes, inflight, n_iter, V, mode, counter, flags, vec, out, work_ns};
                    void* params[] = {&a};
                    CK(cudaLaunchCooperativeKernel((void*)exchange_kernel, dim3(G), dim3(320), params, smem, 0));
                    CK(cudaDeviceSynchronize());
                    unsigned long long h[3];
                    CK(cudaMemcpy(h, out, 24, cudaMemcpyDeviceToHost));
                    const double ns = (double)h[0] / n_iter;
                    printf("V %5d inflight %d work %4d ns  mode %d %-32s : %8.0f ns/iter  (exchange %7.0f ns)  stream %5.0f GB/s  errors %llu\n", V,
                           inflight, work_ns, mode, names[mode], ns, ns - work_ns, (double)h[2] / (double)h[0], h[1]);
                }
            }
        }
    }
    return 0;
}
