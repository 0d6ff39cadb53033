// Row-wise and elementwise kernels (HBM-bound glue between the tensor-core stages).
// Every op keeps the HF bf16 rounding points (SURVEY.md §8a notes); file:line cites are the reference / HF ops replaced.
#include "common.cuh"
#include "kernels.h"

namespace {

// ---------------------------------------------------------------- RMSNorm  (HF:66-71): w * bf16(x * rsqrt(mean(x^2)+eps))
__global__ void rmsnorm_kernel(const bf16* __restrict__ x, int ldx, const bf16* __restrict__ w, bf16* __restrict__ y, int ldy,
                               int rows, int dim, float eps) {
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    int lane = threadIdx.x & 31;
    const bf16* xr = x + (size_t)row * ldx;
    float rstd = rms_rstd_warp(xr, dim, eps, lane);
    bf16* yr = y + (size_t)row * ldy;
    for (int c = lane * 8; c < dim; c += 256)
        *reinterpret_cast<uint4*>(yr + c) = rms_apply8(*reinterpret_cast<const uint4*>(xr + c), *reinterpret_cast<const uint4*>(w + c), rstd);
}

// y[r] = rmsnorm(bf16(low[r / rep] + high[r]))    (padt_decoder.py:220)
__global__ void add_rmsnorm_rep_kernel(const bf16* __restrict__ low, int rep, const bf16* __restrict__ high, const int* __restrict__ high_row_map,
                                       const bf16* __restrict__ w, bf16* __restrict__ y, int rows, int dim, float eps) {
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    int lane = threadIdx.x & 31;
    const bf16* a = low + (size_t)(row / rep) * dim;
    const bf16* b = high + (size_t)(high_row_map ? high_row_map[row] : row) * dim;
    float ss = 0.f;
    for (int c = lane * 2; c < dim; c += 64) {
        float2 av = bf2x_to_f2(*reinterpret_cast<const uint32_t*>(a + c)), bv = bf2x_to_f2(*reinterpret_cast<const uint32_t*>(b + c));
        float s0 = rbf(av.x + bv.x), s1 = rbf(av.y + bv.y);
        ss += s0 * s0 + s1 * s1;
    }
    ss = warp_sum(ss);
    float rstd = 1.0f / sqrtf(ss / (float)dim + eps);
    bf16* yr = y + (size_t)row * dim;
    for (int c = lane * 2; c < dim; c += 64) {
        float2 av = bf2x_to_f2(*reinterpret_cast<const uint32_t*>(a + c)), bv = bf2x_to_f2(*reinterpret_cast<const uint32_t*>(b + c));
        float2 ww = bf2x_to_f2(*reinterpret_cast<const uint32_t*>(w + c));
        float s0 = rbf(av.x + bv.x), s1 = rbf(av.y + bv.y);
        *reinterpret_cast<uint32_t*>(yr + c) = f2_to_bf2x(ww.x * rbf(s0 * rstd), ww.y * rbf(s1 * rstd));
    }
}

// ---------------------------------------------------------------- LayerNorm (vis_norm, padt.py:121,188): fp32 stats, one rounding
__global__ void layernorm_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w, const bf16* __restrict__ b, bf16* __restrict__ y,
                                 int rows, int dim, float eps) {
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    int lane = threadIdx.x & 31;
    const bf16* xr = x + (size_t)row * dim;
    float s = 0.f;
    for (int c = lane; c < dim; c += 32) s += bf2f(xr[c]);
    float mean = warp_sum(s) / (float)dim;
    float v = 0.f;
    for (int c = lane; c < dim; c += 32) { float d = bf2f(xr[c]) - mean; v += d * d; }
    float rstd = 1.0f / sqrtf(warp_sum(v) / (float)dim + eps);
    bf16* yr = y + (size_t)row * dim;
    for (int c = lane; c < dim; c += 32) yr[c] = f2bf((bf2f(xr[c]) - mean) * rstd * bf2f(w[c]) + bf2f(b[c]));
}

__global__ void add_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, bf16* __restrict__ y, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = f2bf(bf2f(a[i]) + bf2f(b[i]));
}
__global__ void add_bcast_row_kernel(const bf16* __restrict__ a, const bf16* __restrict__ r, bf16* __restrict__ y, int rows, int dim) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (long long)rows * dim) y[i] = f2bf(bf2f(a[i]) + bf2f(r[i % dim]));
}
__global__ void sigmoid_kernel(bf16* x, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { float v = bf2f(x[i]); x[i] = f2bf(1.0f / (1.0f + expf(-v))); }
}

// dst[r, :] = src[idx[r], :]   (idx null -> identity); 16-byte chunks
__global__ void gather_rows_kernel(const bf16* __restrict__ src, long long lds, const int* __restrict__ idx, bf16* __restrict__ dst, long long ldd,
                                   int rows, int chunks) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)rows * chunks) return;
    int r = (int)(i / chunks), c = (int)(i % chunks);
    int sr = idx ? idx[r] : r;
    *reinterpret_cast<uint4*>(dst + (size_t)r * ldd + c * 8) = *reinterpret_cast<const uint4*>(src + (size_t)sr * lds + c * 8);
}

// ---------------------------------------------------------------- NeoX rotary in fp32 on bf16 data
// ViT:      HF:156-167  q*cos + rotate_half(q)*sin  with fp32 q,cos,sin, result cast to bf16
// decoder:  padt_decoder.py:39-51 -> flash_attn apply_rotary_emb(x.float(), cos[:, :hd/2], sin[:, :hd/2])
// both evaluate  lo' = lo*c - hi*s ; hi' = hi*c + lo*s  with each product and the sum rounded to fp32 (no FMA).
__global__ void rope_neox_f32_kernel(bf16* __restrict__ x, long long row_stride, int n_heads, int hd, const float* __restrict__ cos,
                                     const float* __restrict__ sin, int cs_ld, const int* __restrict__ cs_row_map, int rows) {
    int half = hd >> 1;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)rows * n_heads * half;
    if (i >= total) return;
    int j = (int)(i % half);
    int h = (int)((i / half) % n_heads);
    int r = (int)(i / ((long long)half * n_heads));
    int cr = cs_row_map ? cs_row_map[r] : r;
    float c = cos[(size_t)cr * cs_ld + j], s = sin[(size_t)cr * cs_ld + j];
    bf16* p = x + (size_t)r * row_stride + (size_t)h * hd;
    float lo = bf2f(p[j]), hi = bf2f(p[j + half]);
    float olo = __fsub_rn(__fmul_rn(lo, c), __fmul_rn(hi, s));
    float ohi = __fadd_rn(__fmul_rn(hi, c), __fmul_rn(lo, s));
    p[j] = f2bf(olo);
    p[j + half] = f2bf(ohi);
}

// same arithmetic, 8 rotation pairs per thread with 16-byte accesses (half % 8 == 0, 16-byte aligned rows)
__global__ void rope_neox_f32_v8_kernel(bf16* __restrict__ x, long long row_stride, int n_heads, int hd, const float* __restrict__ cos,
                                        const float* __restrict__ sin, int cs_ld, const int* __restrict__ cs_row_map, int rows) {
    const int half = hd >> 1, cph = half >> 3;                       // 8-pair chunks per head
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)rows * n_heads * cph;
    if (i >= total) return;
    const int c = (int)(i % cph);
    const int h = (int)((i / cph) % n_heads);
    const int r = (int)(i / ((long long)cph * n_heads));
    const int cr = cs_row_map ? cs_row_map[r] : r;
    const float4* cp = reinterpret_cast<const float4*>(cos + (size_t)cr * cs_ld + c * 8);
    const float4* sp = reinterpret_cast<const float4*>(sin + (size_t)cr * cs_ld + c * 8);
    const float4 c0 = cp[0], c1 = cp[1], s0 = sp[0], s1 = sp[1];
    const float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w}, ss[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    bf16* p = x + (size_t)r * row_stride + (size_t)h * hd + c * 8;
    const uint4 lo4 = *reinterpret_cast<const uint4*>(p), hi4 = *reinterpret_cast<const uint4*>(p + half);
    const uint32_t lo_w[4] = {lo4.x, lo4.y, lo4.z, lo4.w}, hi_w[4] = {hi4.x, hi4.y, hi4.z, hi4.w};
    uint32_t olo[4], ohi[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float2 lo = bf2x_to_f2(lo_w[k]), hi = bf2x_to_f2(hi_w[k]);
        const float a0 = __fsub_rn(__fmul_rn(lo.x, cc[2 * k]), __fmul_rn(hi.x, ss[2 * k]));
        const float a1 = __fsub_rn(__fmul_rn(lo.y, cc[2 * k + 1]), __fmul_rn(hi.y, ss[2 * k + 1]));
        const float b0 = __fadd_rn(__fmul_rn(hi.x, cc[2 * k]), __fmul_rn(lo.x, ss[2 * k]));
        const float b1 = __fadd_rn(__fmul_rn(hi.y, cc[2 * k + 1]), __fmul_rn(lo.y, ss[2 * k + 1]));
        olo[k] = f2_to_bf2x(a0, a1);
        ohi[k] = f2_to_bf2x(b0, b1);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(olo[0], olo[1], olo[2], olo[3]);
    *reinterpret_cast<uint4*>(p + half) = make_uint4(ohi[0], ohi[1], ohi[2], ohi[3]);
}

// cos/sin [rows, 4*quarter] fp32 = cat(rot, rot), rot = [tab[hpos], tab[wpos]]   (HF:405-409, padt.py:73-77)
__global__ void vit_cos_sin_kernel(const float* __restrict__ tab_cos, const float* __restrict__ tab_sin, int quarter, const int* __restrict__ hpos,
                                   const int* __restrict__ wpos, float* __restrict__ cos, float* __restrict__ sin, int rows) {
    int hd = quarter * 4;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)rows * hd) return;
    int r = (int)(i / hd), d = (int)(i % hd);
    int dd = d % (2 * quarter);
    int pos = dd < quarter ? hpos[r] : wpos[r];
    int f = dd % quarter;
    cos[i] = tab_cos[(size_t)pos * quarter + f];
    sin[i] = tab_sin[(size_t)pos * quarter + f];
}

// ---------------------------------------------------------------- LLM M-RoPE in bf16 (HF:659-668) + KV cache write (DynamicCache.update)
// q_embed = bf16(bf16(q*cos) + bf16(rotate_half(q)*sin)); cos/sin are bf16 tables indexed by the t/h/w position plane
// selected per frequency band (mrope_section [s0,s1,s2] doubled over both halves).
__global__ void mrope_kv_kernel(bf16* __restrict__ qkv, int n_tok, int nh, int nkv, int hd, const int* __restrict__ pos3,
                                const bf16* __restrict__ cos_tab, const bf16* __restrict__ sin_tab, int max_pos, int sec0, int sec1,
                                bf16* __restrict__ kcache, bf16* __restrict__ vcache, int B, int L, int s_max) {
    int half = hd >> 1;
    int heads = nh + 2 * nkv;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)n_tok * heads * half;
    if (i >= total) return;
    int j = (int)(i % half);
    int h = (int)((i / half) % heads);
    int tok = (int)(i / ((long long)half * heads));
    bf16* p = qkv + (size_t)tok * heads * hd + (size_t)h * hd;
    int b = tok / L, l = tok % L;
    if (h >= nh + nkv) {  // V head: plain copy into the cache
        int kvh = h - nh - nkv;
        bf16* dst = vcache + (((size_t)b * nkv + kvh) * s_max + l) * hd;
        dst[j] = p[j];
        dst[j + half] = p[j + half];
        return;
    }
    int plane = j < sec0 ? 0 : (j < sec0 + sec1 ? 1 : 2);
    int pos = pos3[(size_t)plane * n_tok + tok];
    pos = pos < 0 ? 0 : (pos >= max_pos ? max_pos - 1 : pos);
    float c = bf2f(cos_tab[(size_t)pos * half + j]), s = bf2f(sin_tab[(size_t)pos * half + j]);
    float lo = bf2f(p[j]), hi = bf2f(p[j + half]);
    float olo = rbf(rbf(lo * c) + rbf(-hi * s));
    float ohi = rbf(rbf(hi * c) + rbf(lo * s));
    if (h < nh) {
        p[j] = f2bf(olo);
        p[j + half] = f2bf(ohi);
    } else {
        int kvh = h - nh;
        bf16* dst = kcache + (((size_t)b * nkv + kvh) * s_max + l) * hd;
        dst[j] = f2bf(olo);
        dst[j + half] = f2bf(ohi);
    }
}

// Same arithmetic, eight frequency pairs per thread with 16-byte accesses (the scalar kernel moved 2 bytes per access and took 26 us per
// prefill layer at L = 2350).  Needs half % 8 == 0 and the mrope section boundaries on multiples of 8 (16 | 24 | 24: yes).
__global__ void mrope_kv_v8_kernel(bf16* __restrict__ qkv, int n_tok, int nh, int nkv, int hd, const int* __restrict__ pos3,
                                   const bf16* __restrict__ cos_tab, const bf16* __restrict__ sin_tab, int max_pos, int sec0, int sec1,
                                   bf16* __restrict__ kcache, bf16* __restrict__ vcache, int B, int L, int s_max) {
    const int half = hd >> 1, chunks = half >> 3;
    const int heads = nh + 2 * nkv;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n_tok * heads * chunks) return;
    const int j = (int)(i % chunks) * 8;
    const int h = (int)((i / chunks) % heads);
    const int tok = (int)(i / ((long long)chunks * heads));
    bf16* p = qkv + (size_t)tok * heads * hd + (size_t)h * hd;
    const int b = tok / L, l = tok % L;
    const uint4 vlo = *reinterpret_cast<const uint4*>(p + j), vhi = *reinterpret_cast<const uint4*>(p + j + half);
    if (h >= nh + nkv) {  // V head: plain copy into the cache
        bf16* dst = vcache + (((size_t)b * nkv + (h - nh - nkv)) * s_max + l) * hd;
        *reinterpret_cast<uint4*>(dst + j) = vlo;
        *reinterpret_cast<uint4*>(dst + j + half) = vhi;
        return;
    }
    const int plane = j < sec0 ? 0 : (j < sec0 + sec1 ? 1 : 2);
    int pos = pos3[(size_t)plane * n_tok + tok];
    pos = pos < 0 ? 0 : (pos >= max_pos ? max_pos - 1 : pos);
    const uint4 vc = *reinterpret_cast<const uint4*>(cos_tab + (size_t)pos * half + j), vs = *reinterpret_cast<const uint4*>(sin_tab + (size_t)pos * half + j);
    const uint32_t lo_w[4] = {vlo.x, vlo.y, vlo.z, vlo.w}, hi_w[4] = {vhi.x, vhi.y, vhi.z, vhi.w};
    const uint32_t c_w[4] = {vc.x, vc.y, vc.z, vc.w}, s_w[4] = {vs.x, vs.y, vs.z, vs.w};
    uint32_t olo[4], ohi[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float2 lo = bf2x_to_f2(lo_w[k]), hi = bf2x_to_f2(hi_w[k]), c = bf2x_to_f2(c_w[k]), sn = bf2x_to_f2(s_w[k]);
        olo[k] = f2_to_bf2x(rbf(rbf(lo.x * c.x) + rbf(-hi.x * sn.x)), rbf(rbf(lo.y * c.y) + rbf(-hi.y * sn.y)));
        ohi[k] = f2_to_bf2x(rbf(rbf(hi.x * c.x) + rbf(lo.x * sn.x)), rbf(rbf(hi.y * c.y) + rbf(lo.y * sn.y)));
    }
    bf16* dst = h < nh ? p : kcache + (((size_t)b * nkv + (h - nh)) * s_max + l) * hd;
    *reinterpret_cast<uint4*>(dst + j) = make_uint4(olo[0], olo[1], olo[2], olo[3]);
    *reinterpret_cast<uint4*>(dst + j + half) = make_uint4(ohi[0], ohi[1], ohi[2], ohi[3]);
}

// ---------------------------------------------------------------- extended embedding gather + image scatter (padt.py:193-219)
__global__ void embed_gather_kernel(const long long* __restrict__ ids, int n_tok, const bf16* __restrict__ embed, int vocab,
                                    const bf16* __restrict__ protos, int n_protos, const bf16* __restrict__ image_embeds,
                                    const int* __restrict__ image_slot, int dim, bf16* __restrict__ out, int* err_flag) {
    int chunks = dim / 8;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n_tok * chunks) return;
    int tok = (int)(i / chunks), c = (int)(i % chunks);
    long long id = ids[tok];
    const bf16* src;
    int slot = image_slot ? image_slot[tok] : -1;
    if (slot >= 0) {
        src = image_embeds + (size_t)slot * dim;
    } else if (id >= 0 && id < vocab) {
        src = embed + (size_t)id * dim;
    } else if (id >= vocab && id < (long long)vocab + n_protos) {
        src = protos + (size_t)(id - vocab) * dim;
    } else {
        if (c == 0 && err_flag) atomicExch(err_flag, 1);   // padt.py:203 assert input_ids.max() < extended rows
        src = embed;
    }
    *reinterpret_cast<uint4*>(out + (size_t)tok * dim + c * 8) = *reinterpret_cast<const uint4*>(src + c * 8);
}

// out[2j] = a[j], out[2j+1] = b[j]  (weight repack: gate/up rows interleaved so one thread owns a (gate,up) pair)
__global__ void interleave_rows_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, bf16* __restrict__ out, int rows, int cols) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)rows * cols) return;
    int r = (int)(i / cols), c = (int)(i % cols);
    out[((size_t)2 * r) * cols + c] = a[i];
    out[((size_t)2 * r + 1) * cols + c] = b[i];
}

}  // namespace

static inline int nblk(long long n, int t) { return (int)((n + t - 1) / t); }

int launch_rmsnorm(const bf16* x, int ldx, const bf16* w, bf16* y, int ldy, int rows, int dim, float eps, cudaStream_t st) {
    if (rows == 0) return 0;
    PADT_REQUIRE(dim % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0, "rmsnorm needs 16-byte rows");
    rmsnorm_kernel<<<cdiv(rows, 8), 256, 0, st>>>(x, ldx, w, y, ldy, rows, dim, eps);
    PADT_LAUNCH_CHECK();
    return 0;
}
int launch_add_rmsnorm_rep_map(const bf16* low, int rep, const bf16* high, const int* high_row_map, const bf16* w, bf16* y, int rows, int dim,
                               float eps, cudaStream_t st) {
    if (rows == 0) return 0;
    PADT_REQUIRE(dim % 2 == 0, "dim");
    add_rmsnorm_rep_kernel<<<cdiv(rows, 8), 256, 0, st>>>(low, rep, high, high_row_map, w, y, rows, dim, eps);
    PADT_LAUNCH_CHECK();
    return 0;
}
int launch_layernorm(const bf16* x, const bf16* w, const bf16* b, bf16* y, int rows, int dim, float eps, cudaStream_t st) {
    if (rows == 0) return 0;
    layernorm_kernel<<<cdiv(rows, 8), 256, 0, st>>>(x, w, b, y, rows, dim, eps);
    PADT_LAUNCH_CHECK();
    return 0;
}
int launch_add(const bf16* a, const bf16* b, bf16* y, long long n, cudaStream_t st) {
    if (n == 0) return 0;
    add_kernel<<<nblk(n, 256), 256, 0, st>>>(a, b, y, n);
    PADT_LAUNCH_CHECK();
    return 0;
}
int launch_add_bcast_row(const bf16* a, const bf16* row, bf16* y, int rows, int dim, cudaStream_t st) {
    if (rows == 0) return 0;
    add_bcast_row_kernel<<<nblk((long long)rows * dim, 256), 256, 0, st>>>(a, row, y, rows, dim);
    PADT_LAUNCH_CHECK();
    return 0;
}
namespace {
__global__ void logit_mask_kernel(bf16* __restrict__ logits, long long ld, int rows_per_sample, int vocab, int n_extra, LogitMaskRanges r, long long total) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // one (row, VRT column) per thread
    if (i >= total) return;
    const long long row = i / n_extra;
    const int c = (int)(i % n_extra);
    const int b = (int)(row / rows_per_sample);
    if (c < r.lo[b] || c >= r.hi[b]) logits[row * ld + vocab + c] = __float2bfloat16_rn(-INFINITY);
}
}  // namespace
int launch_logit_mask(bf16* logits, long long ld, int rows_per_sample, int B, int vocab, int n_ext, const LogitMaskRanges& r, cudaStream_t st) {
    const int n_extra = n_ext - vocab;
    const long long total = (long long)B * rows_per_sample * n_extra;
    if (total <= 0) return 0;
    logit_mask_kernel<<<nblk(total, 256), 256, 0, st>>>(logits, ld, rows_per_sample, vocab, n_extra, r, total);
    PADT_LAUNCH_CHECK();
    return 0;
}
int launch_sigmoid(bf16* x, long long n, cudaStream_t st) {
    if (n == 0) return 0;
    sigmoid_kernel<<<nblk(n, 256), 256, 0, st>>>(x, n);
    PADT_LAUNCH_CHECK();
    return 0;
}
int launch_gather_rows(const bf16* src, int lds, const int* idx, bf16* dst, int ldd, int rows, int dim, cudaStream_t st) {
    if (rows == 0) return 0;
    PADT_REQUIRE(dim % 8 == 0 && lds % 8 == 0 && ldd % 8 == 0, "gather needs 16-byte rows");
    gather_rows_kernel<<<nblk((long long)rows * (dim / 8), 256), 256, 0, st>>>(src, lds, idx, dst, ldd, rows, dim / 8);
    PADT_LAUNCH_CHECK();
    return 0;
}
int launch_copy_rows_strided(const bf16* src, long long lds, bf16* dst, long long ldd, int rows, int dim, cudaStream_t st) {
    if (rows == 0) return 0;
    PADT_REQUIRE(dim % 8 == 0 && lds % 8 == 0 && ldd % 8 == 0, "copy needs 16-byte rows");
    gather_rows_kernel<<<nblk((long long)rows * (dim / 8), 256), 256, 0, st>>>(src, lds, nullptr, dst, ldd, rows, dim / 8);
    PADT_LAUNCH_CHECK();
    return 0;
}
int launch_rope_neox_f32(bf16* x, long long row_stride, int n_heads, int hd, const float* cos, const float* sin, int cs_ld,
                         const int* cs_row_map, int rows, cudaStream_t st) {
    if (rows == 0) return 0;
    const int half = hd / 2;
    if (half % 8 == 0 && hd % 8 == 0 && row_stride % 8 == 0 && cs_ld % 4 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)cos & 15) == 0 &&
        ((uintptr_t)sin & 15) == 0) {
        long long n8 = (long long)rows * n_heads * (half / 8);
        rope_neox_f32_v8_kernel<<<nblk(n8, 256), 256, 0, st>>>(x, row_stride, n_heads, hd, cos, sin, cs_ld, cs_row_map, rows);
        PADT_LAUNCH_CHECK();
        return 0;
    }
    long long n = (long long)rows * n_heads * half;
    rope_neox_f32_kernel<<<nblk(n, 256), 256, 0, st>>>(x, row_stride, n_heads, hd, cos, sin, cs_ld, cs_row_map, rows);
    PADT_LAUNCH_CHECK();
    return 0;
}
int launch_vit_cos_sin(const float* tab_cos, const float* tab_sin, int quarter, const int* hpos, const int* wpos, float* cos, float* sin,
                       int rows, cudaStream_t st) {
    if (rows == 0) return 0;
    vit_cos_sin_kernel<<<nblk((long long)rows * quarter * 4, 256), 256, 0, st>>>(tab_cos, tab_sin, quarter, hpos, wpos, cos, sin, rows);
    PADT_LAUNCH_CHECK();
    return 0;
}
int launch_mrope_kv(bf16* qkv, int n_tok, int nh, int nkv, int hd, const int* pos3, const bf16* cos_tab, const bf16* sin_tab, int max_pos,
                    int sec0, int sec1, bf16* kcache, bf16* vcache, int B, int L, int s_max, cudaStream_t st) {
    if (n_tok == 0) return 0;
    const int half = hd / 2;
    const bool aligned = ((((uintptr_t)qkv) | ((uintptr_t)cos_tab) | ((uintptr_t)sin_tab) | ((uintptr_t)kcache) | ((uintptr_t)vcache)) & 15) == 0;
    if (half % 8 == 0 && sec0 % 8 == 0 && (sec0 + sec1) % 8 == 0 && aligned) {
        const long long n8 = (long long)n_tok * (nh + 2 * nkv) * (half / 8);
        mrope_kv_v8_kernel<<<nblk(n8, 256), 256, 0, st>>>(qkv, n_tok, nh, nkv, hd, pos3, cos_tab, sin_tab, max_pos, sec0, sec1, kcache, vcache, B, L, s_max);
        PADT_LAUNCH_CHECK();
        return 0;
    }
    long long n = (long long)n_tok * (nh + 2 * nkv) * (hd / 2);
    mrope_kv_kernel<<<nblk(n, 256), 256, 0, st>>>(qkv, n_tok, nh, nkv, hd, pos3, cos_tab, sin_tab, max_pos, sec0, sec1, kcache, vcache, B, L, s_max);
    PADT_LAUNCH_CHECK();
    return 0;
}
int launch_embed_gather(const long long* ids, int n_tok, const bf16* embed, int vocab, const bf16* protos, int n_protos,
                        const bf16* image_embeds, const int* image_slot, int dim, bf16* out, int* err_flag, cudaStream_t st) {
    if (n_tok == 0) return 0;
    PADT_REQUIRE(dim % 8 == 0, "dim");
    embed_gather_kernel<<<nblk((long long)n_tok * (dim / 8), 256), 256, 0, st>>>(ids, n_tok, embed, vocab, protos, n_protos, image_embeds,
                                                                                  image_slot, dim, out, err_flag);
    PADT_LAUNCH_CHECK();
    return 0;
}
int launch_interleave_rows(const bf16* a, const bf16* b, bf16* out, int rows, int cols, cudaStream_t st) {
    if (rows == 0) return 0;
    interleave_rows_kernel<<<nblk((long long)rows * cols, 256), 256, 0, st>>>(a, b, out, rows, cols);
    PADT_LAUNCH_CHECK();
    return 0;
}
