// GPU frame preprocessing: the reference's `frame_transform` (mm_utils/utils.py:153-183, used by create_inputs,
// inference.py:69-88) = ToPILImage -> Resize(size, BICUBIC) -> CenterCrop(size) -> ToTensor -> Normalize, bit-exact.
//
// The resize is Pillow's 8-bit ImagingResample (third-party; torchvision's Resize on a PIL image calls Image.resize):
// separable two-pass convolution, horizontal then vertical, with an 8-bit intermediate image; bicubic kernel (a = -0.5)
// of support 2 * max(scale, 1); per output coordinate the taps are normalised in double precision and converted to 22-bit
// fixed point (round half away from zero), the accumulator starts at 1 << 21 and the result is clamp(acc >> 22, 0, 255).
// The coefficient tables are built on the host in double precision with Pillow's exact operation order (they depend only on
// (in_size, out_size) and are cached on the device); the kernels are pure integer multiply-adds, so the output is
// bit-identical to Pillow's. ToTensor / Normalize are the same three IEEE fp32 operations as torchvision
// ((v / 255 - mean) / std, no FMA contraction). HBM-bound: each byte of the frame is read once per pass.
#include "gvl_internal.h"
#include <math.h>
#include <mutex>
#include <vector>

namespace gvl {

namespace {

constexpr int PP_PRECISION_BITS = 32 - 8 - 2;

struct CoeffTable {
    int in_size, out_size, ksize;
    int* bounds;    // device [out_size][2]: first tap, tap count
    int* kk;        // device [out_size][ksize]
};

double bicubic_filter(double x) {
    const double a = -0.5;
    if (x < 0.0) x = -x;
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
    if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0.0;
}

// Pillow Resample.c: precompute_coeffs + normalize_coeffs_8bpc
void build_coeffs(int in_size, int out_size, int& ksize, std::vector<int>& bounds, std::vector<int>& kk) {
    double scale = (double)((float)in_size - 0.0f) / out_size;
    double filterscale = scale;
    if (filterscale < 1.0) filterscale = 1.0;
    const double support = 2.0 * filterscale;
    ksize = (int)ceil(support) * 2 + 1;
    bounds.assign((size_t)out_size * 2, 0);
    kk.assign((size_t)out_size * ksize, 0);
    std::vector<double> k((size_t)ksize);
    const double ss = 1.0 / filterscale;
    for (int xx = 0; xx < out_size; ++xx) {
        const double center = 0.0 + (xx + 0.5) * scale;
        double ww = 0.0;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        for (int x = 0; x < xmax; ++x) {
            const double w = bicubic_filter((x + xmin - center + 0.5) * ss);
            k[x] = w;
            ww += w;
        }
        for (int x = 0; x < xmax; ++x) {
            if (ww != 0.0) k[x] /= ww;
            const double v = k[x];
            kk[(size_t)xx * ksize + x] = v < 0 ? (int)(-0.5 + v * (1 << PP_PRECISION_BITS)) : (int)(0.5 + v * (1 << PP_PRECISION_BITS));
        }
        bounds[(size_t)xx * 2 + 0] = xmin;
        bounds[(size_t)xx * 2 + 1] = xmax;
    }
}

std::mutex g_coeff_mu;
std::vector<CoeffTable> g_coeff_cache;

// tables are tiny (out_size x (2 + ksize) ints) and depend only on the two sizes: built once, kept on the device
int get_coeffs(int in_size, int out_size, CoeffTable* out) {
    std::lock_guard<std::mutex> lock(g_coeff_mu);
    for (const CoeffTable& t : g_coeff_cache)
        if (t.in_size == in_size && t.out_size == out_size) { *out = t; return GVL_OK; }
    CoeffTable t;
    t.in_size = in_size; t.out_size = out_size;
    std::vector<int> bounds, kk;
    build_coeffs(in_size, out_size, t.ksize, bounds, kk);
    if (cudaMalloc(&t.bounds, bounds.size() * sizeof(int)) != cudaSuccess) return GVL_ERR_NOMEM;
    if (cudaMalloc(&t.kk, kk.size() * sizeof(int)) != cudaSuccess) { cudaFree(t.bounds); return GVL_ERR_NOMEM; }
    if (cudaMemcpy(t.bounds, bounds.data(), bounds.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(t.kk, kk.data(), kk.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaFree(t.bounds); cudaFree(t.kk);
        return GVL_ERR_CUDA;
    }
    g_coeff_cache.push_back(t);
    *out = t;
    return GVL_OK;
}

__device__ __forceinline__ unsigned char clip8(int acc) {
    const int v = acc >> PP_PRECISION_BITS;             // arithmetic shift, as Pillow's lookup index
    return (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// horizontal pass: in [planes][h][w] -> out [planes][h][ow]; one thread per output pixel, neighbours share their taps in L1
__global__ void resample_h_kernel(const unsigned char* __restrict__ in, unsigned char* __restrict__ out, int w, int ow,
                                  const int* __restrict__ bounds, const int* __restrict__ kk, int ksize, long long rows) {
    const long long total = rows * ow;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / ow;
        const int xx = (int)(i - row * ow);
        const int xmin = bounds[2 * xx], n = bounds[2 * xx + 1];
        const unsigned char* src = in + row * w + xmin;
        const int* k = kk + (size_t)xx * ksize;
        int acc = 1 << (PP_PRECISION_BITS - 1);
        for (int x = 0; x < n; ++x) acc += (int)src[x] * k[x];
        out[i] = clip8(acc);
    }
}

// vertical pass (optional) + center crop + ToTensor + Normalize: in [planes][h][w] (u8) -> out [planes][size][size] (f32).
// Only the cropped window is produced. VERT = false: the height is already right (pure crop + normalise).
template <bool VERT>
__global__ void resample_v_crop_norm_kernel(const unsigned char* __restrict__ in, float* __restrict__ out, int h, int w,
                                            int size, int top, int left, const int* __restrict__ bounds,
                                            const int* __restrict__ kk, int ksize, int planes, float m0, float m1, float m2,
                                            float s0, float s1, float s2) {
    const long long total = (long long)planes * size * size;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % size);
        const int y = (int)((i / size) % size);
        const int plane = (int)(i / ((long long)size * size));
        const int c = plane % 3;
        const unsigned char* src = in + (size_t)plane * h * w + (x + left);
        unsigned char v;
        if (VERT) {
            const int yy = y + top;
            const int ymin = bounds[2 * yy], n = bounds[2 * yy + 1];
            const int* k = kk + (size_t)yy * ksize;
            int acc = 1 << (PP_PRECISION_BITS - 1);
            for (int t = 0; t < n; ++t) acc += (int)src[(size_t)(ymin + t) * w] * k[t];
            v = clip8(acc);
        } else {
            v = src[(size_t)(y + top) * w];
        }
        const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2);
        const float stdv = c == 0 ? s0 : (c == 1 ? s1 : s2);
        // torchvision: ToTensor = byte -> float32 / 255; Normalize = (t - mean) / std; three separately rounded fp32 operations
        out[i] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)v, 255.0f), mean), stdv);
    }
}

inline int grid_for(long long total) {
    long long g = (total + 255) / 256;
    const long long cap = (long long)num_sms() * 16;
    return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace

}  // namespace gvl

using namespace gvl;

extern "C" {

size_t gvl_frame_transform_workspace(int n, int h, int w, int new_h, int new_w) {
    if (n <= 0 || h <= 0 || w <= 0 || new_h <= 0 || new_w <= 0) return 0;
    return (w != new_w) ? (size_t)n * 3 * h * new_w : 0;     // 8-bit intermediate image of the horizontal pass
}

int gvl_frame_transform(const unsigned char* frames, int n, int h, int w, int new_h, int new_w, int crop_top, int crop_left,
                        int size, const float* mean3, const float* std3, float* out, void* workspace, size_t ws_bytes,
                        void* stream) {
    if (!frames || !out || !mean3 || !std3 || n <= 0 || h <= 0 || w <= 0 || new_h <= 0 || new_w <= 0 || size <= 0) return GVL_ERR_ARG;
    if (crop_top < 0 || crop_left < 0 || crop_top + size > new_h || crop_left + size > new_w) return GVL_ERR_ARG;
    if (ws_bytes < gvl_frame_transform_workspace(n, h, w, new_h, new_w) || (ws_bytes > 0 && !workspace)) return GVL_ERR_ARG;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const int planes = n * 3;
    const unsigned char* cur = frames;
    int cur_w = w;
    if (w != new_w) {
        CoeffTable th;
        int rc = get_coeffs(w, new_w, &th);
        if (rc != GVL_OK) return rc;
        const long long rows = (long long)planes * h;
        resample_h_kernel<<<grid_for(rows * new_w), 256, 0, s>>>(frames, reinterpret_cast<unsigned char*>(workspace), w, new_w,
                                                                th.bounds, th.kk, th.ksize, rows);
        g_launch_count++;
        if (cudaGetLastError() != cudaSuccess) return GVL_ERR_CUDA;
        cur = reinterpret_cast<const unsigned char*>(workspace);
        cur_w = new_w;
    }
    const long long total = (long long)planes * size * size;
    if (h != new_h) {
        CoeffTable tv;
        int rc = get_coeffs(h, new_h, &tv);
        if (rc != GVL_OK) return rc;
        resample_v_crop_norm_kernel<true><<<grid_for(total), 256, 0, s>>>(cur, out, h, cur_w, size, crop_top, crop_left, tv.bounds,
                                                                           tv.kk, tv.ksize, planes, mean3[0], mean3[1], mean3[2],
                                                                           std3[0], std3[1], std3[2]);
    } else {
        resample_v_crop_norm_kernel<false><<<grid_for(total), 256, 0, s>>>(cur, out, h, cur_w, size, crop_top, crop_left, nullptr,
                                                                            nullptr, 0, planes, mean3[0], mean3[1], mean3[2],
                                                                            std3[0], std3[1], std3[2]);
    }
    g_launch_count++;
    return cudaGetLastError() == cudaSuccess ? GVL_OK : GVL_ERR_CUDA;
}

}  // extern "C"
