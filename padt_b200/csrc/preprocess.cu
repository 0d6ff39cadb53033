// Image pre-processing on the GPU: what the reference's host-side processor call computes for every request
// (eval/test_demo.py:87-93, eval/evaluation_scripts/utils.py:221-229 -> transformers' Qwen2-VL image processor, PIL backend in 4.50):
//   PIL Image.resize(BICUBIC) on uint8 RGB -> * 1/255 (float64 product, float32 cast) -> (x - mean) / std (float32) -> patch rows
//   [grid_h * grid_w, 3 * 2 * 14 * 14] in (h/2, w/2, 2, 2) row order with the frame repeated twice in time.
// Bit-exact with Pillow's integer resampler (src/libImaging/Resample.c): coefficients in IEEE double without contraction, 22-bit fixed
// point, int32 accumulation from 1 << 21, arithmetic shift, clip to 0..255; horizontal pass first, vertical pass over the uint8
// intermediate.  The vertical pass, rescale, normalise, the bf16 cast of the patch-embed input (HF:108-114 casts to the conv dtype) and
// the patch layout are one kernel: the resized image is never materialised.  HBM-bound byte work: coalesced rows, no tensor cores.
#include <cmath>
#include <cfenv>
#include "common.cuh"
#include "kernels.h"
#include "../../include/padt_b200.h"

namespace {

constexpr int PRECISION_BITS = 32 - 8 - 2;

// Resample.c bicubic_filter (a = -0.5), every operation rounded separately (the host compilers Pillow ships with do not contract on x86-64)
__device__ __forceinline__ double bicubic_filter(double x) {
    if (x < 0.0) x = -x;
    if (x < 1.0) {
        double t = __dmul_rn(1.5, x);
        t = __dsub_rn(t, 2.5);
        t = __dmul_rn(t, x);
        t = __dmul_rn(t, x);
        return __dadd_rn(t, 1.0);
    }
    if (x < 2.0) {
        double t = __dsub_rn(x, 5.0);
        t = __dmul_rn(t, x);
        t = __dadd_rn(t, 8.0);
        t = __dmul_rn(t, x);
        t = __dsub_rn(t, 4.0);
        return __dmul_rn(t, -0.5);
    }
    return 0.0;
}

// Resample.c precompute_coeffs + normalize_coeffs_8bpc for the box (0, in_size): one thread per output index
__global__ void resample_coeffs_kernel(int in_size, int out_size, int ksize, int* __restrict__ bounds, int* __restrict__ kk) {
    const int xx = blockIdx.x * blockDim.x + threadIdx.x;
    if (xx >= out_size) return;
    const double scale = __ddiv_rn((double)in_size, (double)out_size);
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = __dmul_rn(2.0, filterscale);
    const double ss = __ddiv_rn(1.0, filterscale);
    const double center = __dmul_rn((double)xx + 0.5, scale);
    int xmin = __double2int_rz(__dadd_rn(__dsub_rn(center, support), 0.5));
    if (xmin < 0) xmin = 0;
    int xmax = __double2int_rz(__dadd_rn(__dadd_rn(center, support), 0.5));
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x)
        ww = __dadd_rn(ww, bicubic_filter(__dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss)));
    int* k = kk + (size_t)xx * ksize;
    for (int x = 0; x < xmax; ++x) {
        double w = bicubic_filter(__dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss));
        if (ww != 0.0) w = __ddiv_rn(w, ww);
        const double s = __dmul_rn(w, (double)(1 << PRECISION_BITS));
        k[x] = w < 0 ? __double2int_rz(__dadd_rn(-0.5, s)) : __double2int_rz(__dadd_rn(0.5, s));
    }
    for (int x = xmax; x < ksize; ++x) k[x] = 0;
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
}

__device__ __forceinline__ int clip8(int acc) {
    const int v = acc >> PRECISION_BITS;
    return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// horizontal pass: in [H, W, 3] uint8 -> tmp [H, out_w, 3] uint8 (ImagingResampleHorizontal_8bpc)
__global__ void resample_h_kernel(const uint8_t* __restrict__ in, int H, int W, int out_w, int ksize, const int* __restrict__ bounds,
                                  const int* __restrict__ kk, uint8_t* __restrict__ tmp) {
    const int ox = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (ox >= out_w) return;
    const int xmin = bounds[2 * ox], xmax = bounds[2 * ox + 1];
    const int* k = kk + (size_t)ox * ksize;
    const uint8_t* row = in + ((size_t)y * W + xmin) * 3;
    int a0 = 1 << (PRECISION_BITS - 1), a1 = a0, a2 = a0;
    for (int x = 0; x < xmax; ++x) {
        const int c = k[x];
        a0 += row[3 * x] * c; a1 += row[3 * x + 1] * c; a2 += row[3 * x + 2] * c;
    }
    uint8_t* o = tmp + ((size_t)y * out_w + ox) * 3;
    o[0] = (uint8_t)clip8(a0); o[1] = (uint8_t)clip8(a1); o[2] = (uint8_t)clip8(a2);
}

struct NormParams { float mean[3], stdv[3]; double rescale; };

// vertical pass (ImagingResampleVertical_8bpc) fused with rescale, normalise, the output cast and the patch layout
template <typename OutT>
__global__ void resample_v_patchify_kernel(const uint8_t* __restrict__ tmp, int out_h, int out_w, int ksize, const int* __restrict__ bounds,
                                           const int* __restrict__ kk, NormParams np, int patch, int temporal, int merge, OutT* __restrict__ pv) {
    const int ox = blockIdx.x * blockDim.x + threadIdx.x, oy = blockIdx.y;
    if (ox >= out_w) return;
    const int ymin = bounds[2 * oy], ymax = bounds[2 * oy + 1];
    const int* k = kk + (size_t)oy * ksize;
    int a[3] = {1 << (PRECISION_BITS - 1), 1 << (PRECISION_BITS - 1), 1 << (PRECISION_BITS - 1)};
    for (int y = 0; y < ymax; ++y) {
        const uint8_t* px = tmp + ((size_t)(ymin + y) * out_w + ox) * 3;
        const int c = k[y];
        a[0] += px[0] * c; a[1] += px[1] * c; a[2] += px[2] * c;
    }
    // patch row / column of this pixel: rows in (h / merge, w / merge, merge, merge) order, columns (channel, time, py, px)
    const int gw = out_w / patch;
    const int ph = oy / patch, pw = ox / patch, py = oy % patch, px = ox % patch;
    const size_t prow = ((size_t)(ph / merge) * (gw / merge) + (pw / merge)) * (merge * merge) + (ph % merge) * merge + (pw % merge);
    const int pp = patch * patch, row_len = 3 * temporal * pp;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float x32 = __double2float_rn(__dmul_rn((double)clip8(a[c]), np.rescale));
        const float y = __fdiv_rn(__fsub_rn(x32, np.mean[c]), np.stdv[c]);
        OutT* dst = pv + prow * row_len + (size_t)c * temporal * pp + py * patch + px;
        for (int t = 0; t < temporal; ++t) {
            if constexpr (sizeof(OutT) == 2) dst[(size_t)t * pp] = __float2bfloat16_rn(y);
            else dst[(size_t)t * pp] = y;
        }
    }
}

size_t align256(size_t x) { return (x + 255) & ~size_t(255); }
int ksize_of(int in_size, int out_size) {
    double scale = (double)in_size / out_size;
    if (scale < 1.0) scale = 1.0;
    return (int)ceil(2.0 * scale) * 2 + 1;
}

}  // namespace

extern "C" int padt_smart_resize(int height, int width, int factor, int64_t min_pixels, int64_t max_pixels, int32_t* out_hw) {
    PADT_REQUIRE(out_hw && height > 0 && width > 0 && factor > 0, "arguments");
    const int lo = height < width ? height : width, hi = height < width ? width : height;
    PADT_REQUIRE((double)hi / lo <= 200.0, "absolute aspect ratio must be smaller than 200");
    // Python round() is round-half-to-even = nearbyint under the default rounding mode
    const int saved = fegetround();
    fesetround(FE_TONEAREST);
    long long h_bar = (long long)nearbyint((double)height / factor) * factor;
    long long w_bar = (long long)nearbyint((double)width / factor) * factor;
    fesetround(saved);
    if (h_bar * w_bar > max_pixels) {
        const double beta = sqrt(((double)height * width) / (double)max_pixels);
        h_bar = (long long)floor((double)height / beta / factor) * factor;
        w_bar = (long long)floor((double)width / beta / factor) * factor;
        if (h_bar < factor) h_bar = factor;
        if (w_bar < factor) w_bar = factor;
    } else if (h_bar * w_bar < min_pixels) {
        const double beta = sqrt((double)min_pixels / ((double)height * width));
        h_bar = (long long)ceil((double)height * beta / factor) * factor;
        w_bar = (long long)ceil((double)width * beta / factor) * factor;
    }
    out_hw[0] = (int32_t)h_bar; out_hw[1] = (int32_t)w_bar;
    return 0;
}

extern "C" size_t padt_preprocess_workspace_bytes(int H, int W, int out_h, int out_w) {
    if (H <= 0 || W <= 0 || out_h <= 0 || out_w <= 0) return 0;
    const size_t kh = ksize_of(W, out_w), kv = ksize_of(H, out_h);
    return align256((size_t)out_w * kh * 4) + align256((size_t)out_w * 8) + align256((size_t)out_h * kv * 4) + align256((size_t)out_h * 8) +
           align256((size_t)H * out_w * 3);
}

extern "C" int padt_preprocess_image(const uint8_t* rgb, int H, int W, int out_h, int out_w, const float* mean3_host, const float* std3_host,
                                     double rescale_factor, int patch, int temporal, int merge, void* pixel_values, int out_f32, void* workspace,
                                     size_t workspace_bytes, padt_stream_t stream) {
    PADT_REQUIRE(rgb && pixel_values && workspace && mean3_host && std3_host, "null pointer");
    PADT_REQUIRE(H > 0 && W > 0 && out_h > 0 && out_w > 0 && patch > 0 && temporal > 0 && merge > 0, "sizes");
    PADT_REQUIRE(out_h % (patch * merge) == 0 && out_w % (patch * merge) == 0, "output size must be a multiple of patch * merge (smart_resize)");
    PADT_REQUIRE(H < (1 << 24) && W < (1 << 24), "image side too large");
    PADT_REQUIRE(workspace_bytes >= padt_preprocess_workspace_bytes(H, W, out_h, out_w), "workspace too small (padt_preprocess_workspace_bytes)");
    cudaStream_t st = (cudaStream_t)stream;
    const int kh = ksize_of(W, out_w), kv = ksize_of(H, out_h);
    uint8_t* w8 = static_cast<uint8_t*>(workspace);
    int* kk_h = reinterpret_cast<int*>(w8); w8 += align256((size_t)out_w * kh * 4);
    int* b_h = reinterpret_cast<int*>(w8); w8 += align256((size_t)out_w * 8);
    int* kk_v = reinterpret_cast<int*>(w8); w8 += align256((size_t)out_h * kv * 4);
    int* b_v = reinterpret_cast<int*>(w8); w8 += align256((size_t)out_h * 8);
    uint8_t* tmp = w8;
    resample_coeffs_kernel<<<cdiv(out_w, 128), 128, 0, st>>>(W, out_w, kh, b_h, kk_h);
    PADT_LAUNCH_CHECK();
    resample_coeffs_kernel<<<cdiv(out_h, 128), 128, 0, st>>>(H, out_h, kv, b_v, kk_v);
    PADT_LAUNCH_CHECK();
    resample_h_kernel<<<dim3(cdiv(out_w, 128), H), 128, 0, st>>>(rgb, H, W, out_w, kh, b_h, kk_h, tmp);
    PADT_LAUNCH_CHECK();
    NormParams np;
    for (int c = 0; c < 3; ++c) { np.mean[c] = mean3_host[c]; np.stdv[c] = std3_host[c]; }
    np.rescale = rescale_factor;
    if (out_f32)
        resample_v_patchify_kernel<float><<<dim3(cdiv(out_w, 128), out_h), 128, 0, st>>>(tmp, out_h, out_w, kv, b_v, kk_v, np, patch, temporal, merge,
                                                                                        static_cast<float*>(pixel_values));
    else
        resample_v_patchify_kernel<bf16><<<dim3(cdiv(out_w, 128), out_h), 128, 0, st>>>(tmp, out_h, out_w, kv, b_v, kk_v, np, patch, temporal, merge,
                                                                                       static_cast<bf16*>(pixel_values));
    PADT_LAUNCH_CHECK();
    g_padt_launches += 4;
    return 0;
}
