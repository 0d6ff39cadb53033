// PaDT-decoder-specific kernels: object query assembly and the fused mask head + scatter.
#include "common.cuh"
#include "kernels.h"

namespace {

// padt_decoder.py:199-206: per object [box_tok, score_tok, mask_tok, proj(vrt_k) + vp_embedding ...]
// src[r] = -1,-2,-3 for the three learned tokens, else the row of the projected VRT feature.
__global__ void build_queries_kernel(const bf16* __restrict__ tokens3, const bf16* __restrict__ vp_proj, const bf16* __restrict__ vp_emb,
                                     const int* __restrict__ src, int rows, int dim, bf16* __restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)rows * dim) return;
    int r = (int)(i / dim), c = (int)(i % dim);
    int s = src[r];
    out[i] = s < 0 ? tokens3[(size_t)(-s - 1) * dim + c] : f2bf(bf2f(vp_proj[(size_t)s * dim + c]) + bf2f(vp_emb[c]));
}

// padt_decoder.py:240-274.  `sub` [4N, 4*ch]: row (n,a,b), col (c,d,k) = GELU(upscaling2(...)); logit of sub-pixel
// (y,x) = (2a+c, 2b+d) of patch n is bf16(sum_k bf16(sub*hyper[obj(n)][k])); patch n of object o lands at
// (pos // W_o, pos % W_o), pos = n - cu_patch[o]  (the reference treats the window-ordered index as raster: kept).
__global__ void mask_head_kernel(const bf16* __restrict__ sub, const bf16* __restrict__ hyper, const int* __restrict__ cu_patch,
                                 const int* __restrict__ obj_w, int n_obj, int ch, int hmax, int wmax, bf16* __restrict__ masks, long long total) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int cd = (int)(i & 3);
    int ab = (int)((i >> 2) & 3);
    int n = (int)(i >> 4);
    int lo = 0, hi = n_obj - 1;   // largest o with cu_patch[o] <= n
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (cu_patch[mid] <= n) lo = mid; else hi = mid - 1;
    }
    int o = lo;
    int pos = n - cu_patch[o];
    int W = obj_w[o];
    int row = pos / W, col = pos % W;
    const bf16* s = sub + ((size_t)n * 4 + ab) * (4 * ch) + (size_t)cd * ch;
    const bf16* hv = hyper + (size_t)o * ch;
    float acc = 0.f;
    for (int k = 0; k < ch; ++k) acc += rbf(bf2f(s[k]) * bf2f(hv[k]));
    int a = ab >> 1, b = ab & 1, c = cd >> 1, d = cd & 1;
    int y = row * 4 + 2 * a + c, x = col * 4 + 2 * b + d;
    masks[((size_t)o * (4 * hmax) + y) * (4 * wmax) + x] = f2bf(acc);
}

__global__ void gelu_kernel(bf16* x, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = f2bf(gelu_erf_f(bf2f(x[i])));
}

}  // namespace

int launch_build_queries(const bf16* tokens3, const bf16* vp_proj, const bf16* vp_emb, const int* src, int rows, int dim, bf16* out,
                         cudaStream_t st) {
    if (rows == 0) return 0;
    long long n = (long long)rows * dim;
    build_queries_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(tokens3, vp_proj, vp_emb, src, rows, dim, out);
    PADT_LAUNCH_CHECK();
    return 0;
}

int launch_mask_head(const bf16* sub, const bf16* hyper, const int* cu_patch, const int* obj_w, int n_obj, int n_patches, int ch, int hmax,
                     int wmax, bf16* masks, cudaStream_t st) {
    if (n_patches == 0) return 0;
    long long total = (long long)n_patches * 16;
    mask_head_kernel<<<(int)((total + 255) / 256), 256, 0, st>>>(sub, hyper, cu_patch, obj_w, n_obj, ch, hmax, wmax, masks, total);
    PADT_LAUNCH_CHECK();
    return 0;
}

int launch_gelu(bf16* x, long long n, cudaStream_t st) {
    if (n == 0) return 0;
    gelu_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(x, n);
    PADT_LAUNCH_CHECK();
    return 0;
}
