// Micro-benchmark: MUFU.EX2 throughput per SM on this GPU (softmax bound of the attention kernel), alone and mixed with the FMA-pipe
// work of a softmax inner loop, for 1 / 2 / 4 warps per scheduler.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe_mufu tools/probe_mufu.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// mode 0: 8 independent ex2 per iteration; mode 1: softmax-like: fma + ex2 + add (+ pack every 2); mode 2: polynomial exp2 on the FMA pipe only
template <int MODE>
__global__ void k(float* out, int iters, float seed, long long* cyc) {
    float v[8], acc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = seed * (i + 1) - threadIdx.x * 1e-3f;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) v[i] = ex2(v[i]) - 1.0f;
            if (MODE == 1) { const float p = ex2(fmaf(v[i], 0.125f, -0.5f)); acc += p; v[i] = p - 1.0f; }
            if (MODE == 2) {
                // Cody-Waite: 2^x = 2^floor(x) * P3(frac); magic-number floor, exponent add by integer shift
                float x = fmaf(v[i], 0.125f, -0.5f);
                x = fmaxf(x, -126.f);
                const float r = x + 12582912.f;                  // 1.5 * 2^23: round to nearest integer
                const float fl = r - 12582912.f;
                const float f = x - fl;                          // in [-0.5, 0.5]
                float p = fmaf(f, 0.0555041f, 0.2402265f);
                p = fmaf(f, p, 0.6931472f);
                p = fmaf(f, p, 1.0f);
                const float y = __int_as_float(__float_as_int(p) + (__float_as_int(r) << 23));
                acc += y;
                v[i] = y - 1.0f;
            }
        }
    }
    const long long t1 = clock64();
    float s = acc;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int G = prop.multiProcessorCount;
    float* out;
    long long* cyc;
    CK(cudaMalloc(&out, (size_t)G * 1024 * 4));
    CK(cudaMalloc(&cyc, 8));
    const int iters = 4096;
    const char* names[] = {"ex2 only", "fma + ex2 + add", "FMA-pipe polynomial exp2"};
    for (int mode = 0; mode < 3; ++mode)
        for (int threads : {128, 256, 512, 1024}) {
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) k<0><<<G, threads>>>(out, iters, 0.37f, cyc);
                if (mode == 1) k<1><<<G, threads>>>(out, iters, 0.37f, cyc);
                if (mode == 2) k<2><<<G, threads>>>(out, iters, 0.37f, cyc);
                CK(cudaDeviceSynchronize());
            }
            long long c;
            CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
            const double ops = (double)iters * 8 * threads;
            printf("%-28s %4d threads/SM (%d warps/scheduler): %8lld clk  -> %.2f exp2 / clk / SM\n", names[mode], threads, threads / 128, c, ops / c);
        }
    return 0;
}
