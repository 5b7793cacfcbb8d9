// Micro-benchmark: legacy mma.sync (HMMA.16816 bf16) issue rate on this GPU -- is the decode step's GEMV consumer loop bound by it?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe_hmma tools/probe_hmma.cu && tools/probe_hmma
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int CHAINS>
__global__ void k_mma(float* out, int iters, uint32_t seed, long long* cyc) {
    float acc[CHAINS][4];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    uint32_t a = seed + threadIdx.x, b = seed * 3 + threadIdx.x;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) mma16816(acc[i], a, a, a + i, a + i, b, b + i);
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// the FMA-pipe alternative for a GEMV item: 8 bf16 weights x 8 bf16 activations per 128-bit pair -> unpack + FFMA
__global__ void k_fma(const uint4* __restrict__ w, const uint4* __restrict__ x, float* out, int iters, long long* cyc) {
    __shared__ uint4 sw[256 * 4], sx[256];
    for (int i = threadIdx.x; i < 256 * 4; i += blockDim.x) sw[i] = w[i];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sx[i] = x[i];
    __syncthreads();
    float acc0 = 0.f, acc1 = 0.f;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint4 a = sw[((threadIdx.x + it) & 255) * 4 + u], b = sx[(threadIdx.x + u + it) & 255];
            const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                acc0 = fmaf(__uint_as_float(aw[j] << 16), __uint_as_float(bw[j] << 16), acc0);
                acc1 = fmaf(__uint_as_float(aw[j] & 0xffff0000u), __uint_as_float(bw[j] & 0xffff0000u), acc1);
            }
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc0 + acc1;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int G = prop.multiProcessorCount;
    float* out;
    long long* cyc;
    uint4 *w, *x;
    CK(cudaMalloc(&out, (size_t)G * 1024 * 4));
    CK(cudaMalloc(&cyc, 8));
    CK(cudaMalloc(&w, 256 * 4 * 16));
    CK(cudaMalloc(&x, 256 * 16));
    CK(cudaMemset(w, 0x3c, 256 * 4 * 16));
    CK(cudaMemset(x, 0x3c, 256 * 16));
    const int iters = 4096;
    for (int threads : {128, 256, 512}) {
        for (int rep = 0; rep < 2; ++rep) { k_mma<4><<<G, threads>>>(out, iters, 1u, cyc); CK(cudaDeviceSynchronize()); }
        long long c;
        CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
        const double n = (double)iters * 4 * (threads / 32);
        printf("mma.sync m16n8k16 bf16, 4 independent accumulators, %d warps/SM (%d per scheduler): %lld clk -> %.1f clk per HMMA per scheduler, %.0f FLOP/clk/SM\n",
               threads / 32, threads / 128, c, (double)c / ((double)iters * 4 * (threads / 128)), n * 4096.0 / c);
    }
    for (int threads : {256, 512}) {
        for (int rep = 0; rep < 2; ++rep) { k_fma<<<G, threads>>>(w, x, out, iters, cyc); CK(cudaDeviceSynchronize()); }
        long long c;
        CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
        const double macs = (double)iters * 32 * threads;
        printf("unpack + FFMA GEMV inner loop from shared memory, %d warps/SM: %lld clk -> %.1f MAC/clk/SM = %.1f weight bytes/clk/SM\n", threads / 32, c,
               macs / c, macs * 2 / c);
    }
    return 0;
}
