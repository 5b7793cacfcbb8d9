// Micro-benchmark: how fast can 148 persistent CTAs stream HBM through a shared-memory ring with cp.async.bulk, as a
// function of the copy size, and how does that compare with plain LDG streaming?  (decode_mega.cu design input)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_stream tools/probe_stream.cu && ./probe_stream
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
                 : "memory");
}

constexpr int STAGE_BYTES = 65536;

// pattern 0: every stage is one contiguous 64 KB block, split into `chunk`-byte copies
// pattern 1: decode-like: stage = 8 items x 8 rows x 1 KB, rows 6 KB apart (K = 3072), items consecutive k-segments
__global__ void __launch_bounds__(288, 1) ring_kernel(const uint8_t* __restrict__ buf, size_t n_blocks, int chunk, int stages,
                                                      int pattern, int touch, unsigned long long* sink) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full[4], empty[4];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < 4; ++s) { mbar_init(smem_u32(&full[s]), 1); mbar_init(smem_u32(&empty[s]), 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    int stage = 0;
    uint32_t phase = 0;
    const size_t G = gridDim.x;
    if (warp == 8) {
        const int blk = pattern == 1 ? 49152 : STAGE_BYTES;
        const int per = blk / chunk;
        for (size_t b = blockIdx.x; b < n_blocks; b += G) {
            mbar_wait(smem_u32(&empty[stage]), phase ^ 1);
            const uint32_t bar = smem_u32(&full[stage]);
            if (lane == 0) mbar_expect_tx(bar, blk);
            __syncwarp();
            const uint8_t* src = buf + b * blk;
            for (int i = lane; i < per; i += 32) {
                size_t off = (size_t)i * chunk;
                if (pattern == 1) {
                    // 48 copies of 1 KB: k segment = i / 8, row = i % 8; the block is 8 rows x 6 KB (K = 3072), contiguous
                    const int seg = i >> 3, row = i & 7;
                    off = (size_t)row * 6144 + (size_t)seg * 1024;
                }
                bulk_g2s(smem_u32(smem) + stage * STAGE_BYTES + i * chunk, src + off, chunk, bar);
            }
            if (++stage == stages) { stage = 0; phase ^= 1; }
        }
        return;
    }
    unsigned long long acc = 0;
    for (size_t b = blockIdx.x; b < n_blocks; b += G) {
        mbar_wait(smem_u32(&full[stage]), phase);
        if (touch) {
            const uint4* p = reinterpret_cast<const uint4*>(smem + (size_t)stage * STAGE_BYTES + warp * 8192);
            for (int i = lane; i < 512; i += 32) { const uint4 v = p[i]; acc += v.x ^ v.y ^ v.z ^ v.w; }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&empty[stage]));
        if (++stage == stages) { stage = 0; phase ^= 1; }
    }
    if (acc == 0x123456789abcdefull) *sink = acc;
}

// plain LDG streaming: every warp reads 512 B per instruction, `unroll` independent loads in flight
template <int U>
__global__ void __launch_bounds__(256) ldg_kernel(const uint4* __restrict__ buf, size_t n_vec, unsigned long long* sink) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    unsigned long long acc = 0;
    size_t i = tid;
    for (; i + (U - 1) * nth < n_vec; i += U * nth) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(buf + i + u * nth));
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    if (acc == 0x123456789abcdefull) *sink = acc;
}

int main(int argc, char** argv) {
    const size_t bytes = (size_t)4 << 30;
    uint8_t* buf;
    unsigned long long* sink;
    CK(cudaMalloc(&buf, bytes + (1 << 20)));
    CK(cudaMalloc(&sink, 8));
    CK(cudaMemset(buf, 1, bytes + (1 << 20)));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int G = prop.multiProcessorCount;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaFuncSetAttribute(ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * STAGE_BYTES + 1024));
    const size_t n_blocks = bytes / STAGE_BYTES;
    struct Cfg { int chunk, stages, pattern, touch; };
    const Cfg cfgs[] = {{1024, 3, 0, 0}, {1024, 3, 1, 0}, {2048, 3, 0, 0}, {4096, 3, 0, 0}, {8192, 3, 0, 0}, {16384, 3, 0, 0},
                        {65536, 3, 0, 0}, {1024, 2, 0, 0}, {8192, 2, 0, 0}, {1024, 3, 1, 1}, {8192, 3, 0, 1}};
    for (const Cfg& c : cfgs) {
        float best = 1e9f;
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaEventRecord(e0));
            ring_kernel<<<G, 288, c.stages * STAGE_BYTES>>>(buf, n_blocks, c.chunk, c.stages, c.pattern, c.touch, sink);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            best = ms < best ? ms : best;
        }
        CK(cudaGetLastError());
        const double moved = c.pattern == 1 ? (double)n_blocks * 49152 : (double)bytes;
        printf("ring   chunk %6d stages %d pattern %d touch %d : %.3f ms  %.0f GB/s\n", c.chunk, c.stages, c.pattern, c.touch, best,
               moved / best / 1e6);
    }
    const size_t n_vec = bytes / 16;
    for (int mult = 1; mult <= 8; mult *= 2) {
        float best = 1e9f;
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaEventRecord(e0));
            ldg_kernel<8><<<G * mult, 256>>>(reinterpret_cast<const uint4*>(buf), n_vec, sink);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            best = ms < best ? ms : best;
        }
        printf("ldg    U=8 ctas/SM %d : %.3f ms  %.0f GB/s\n", mult, best, bytes / best / 1e6);
    }
    for (int mult = 2; mult <= 8; mult *= 2) {
        float best = 1e9f;
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaEventRecord(e0));
            ldg_kernel<16><<<G * mult, 256>>>(reinterpret_cast<const uint4*>(buf), n_vec, sink);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            best = ms < best ? ms : best;
        }
        printf("ldg    U=16 ctas/SM %d : %.3f ms  %.0f GB/s\n", mult, best, bytes / best / 1e6);
    }
    return 0;
}
