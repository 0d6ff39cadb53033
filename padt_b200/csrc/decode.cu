// Autoregressive decode step (one new token per row), HBM-bound weight streaming.
//
// Replaces, per generated token, the reference's forward_main past_image_embeds branch (padt.py:221-229,268-301):
// ext[input_ids] gather without materialising cat(embed_tokens, prototypes) (padt.py:226-229), 36/28 decoder layers
// (HF:762-828) at q_len 1 with KV-cache append, final norm, VRT-extended logits + per-row mask (padt.py:292-301) and
// the greedy token selection / finished-row bookkeeping of _sample_vision_token (padt.py:713,745-757).
//
// v1 layout: 5 kernels per layer + sample/embed + final-norm/logits, all step-dependent values read from a
// device-resident state block so the whole step is one replayable CUDA graph (no host sync per token).
#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int GEMV_THREADS = 256;
constexpr int GEMV_WARPS = GEMV_THREADS / 32;

enum { PRO_NONE = 0, PRO_RMSNORM = 1 };
enum { GE_BIAS = 0, GE_RESID = 1, GE_SWIGLU = 2, GE_LOGITS = 3 };

struct GemvArgs {
    const bf16* x;        // PRO_NONE: [B,K] input ; PRO_RMSNORM: residual stream h [B,K]
    const bf16* ln_w;     // PRO_RMSNORM
    float eps;
    const bf16* W;        // [N,K] (GE_LOGITS: head rows [vocab,K])
    const bf16* W2;       // GE_LOGITS: prototype rows [n_protos,K]
    const bf16* bias;     // GE_BIAS (nullable)
    bf16* y;              // GE_BIAS: [B,N]; GE_RESID: h [B,N] updated in place; GE_SWIGLU: [B,N/2]
    int B, N, K, vocab;
    // GE_LOGITS
    const int* state;     // device state block
    float* amax_val; int* amax_idx;
    bf16* last_hidden;    // PRO_RMSNORM + GE_LOGITS: normalised rows also stored here
    bf16* hidden_steps;   // optional log [T,B,K]
    int max_steps;
};

// state block layout: see kernels.h
constexpr int ST_HDR = PADT_STATE_HDR, ST_BMAX = PADT_STATE_BMAX;
__device__ __host__ inline int st_kvstart(int b) { return ST_HDR + 0 * ST_BMAX + b; }
__device__ __host__ inline int st_delta(int b) { return ST_HDR + 1 * ST_BMAX + b; }
__device__ __host__ inline int st_vlo(int b) { return ST_HDR + 2 * ST_BMAX + b; }
__device__ __host__ inline int st_vhi(int b) { return ST_HDR + 3 * ST_BMAX + b; }
__device__ __host__ inline int st_unfin(int b) { return ST_HDR + 4 * ST_BMAX + b; }

template <int NB>
__device__ __forceinline__ void dot8(const uint4& w, const bf16* xs, int K, int c, float* acc) {
    float2 w0 = bf2x_to_f2(w.x), w1 = bf2x_to_f2(w.y), w2 = bf2x_to_f2(w.z), w3 = bf2x_to_f2(w.w);
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        uint4 xv = *reinterpret_cast<const uint4*>(xs + (size_t)b * K + c * 8);
        float2 x0 = bf2x_to_f2(xv.x), x1 = bf2x_to_f2(xv.y), x2 = bf2x_to_f2(xv.z), x3 = bf2x_to_f2(xv.w);
        float a = acc[b];
        a = fmaf(w0.x, x0.x, a); a = fmaf(w0.y, x0.y, a);
        a = fmaf(w1.x, x1.x, a); a = fmaf(w1.y, x1.y, a);
        a = fmaf(w2.x, x2.x, a); a = fmaf(w2.y, x2.y, a);
        a = fmaf(w3.x, x3.x, a); a = fmaf(w3.y, x3.y, a);
        acc[b] = a;
    }
}

// One warp computes two weight rows against NB activation rows held in shared memory.
template <int NB>
__device__ __forceinline__ void warp_rows2(const bf16* __restrict__ r0, const bf16* __restrict__ r1, int K, const bf16* xs, int lane,
                                           float* acc0, float* acc1) {
#pragma unroll
    for (int b = 0; b < NB; ++b) { acc0[b] = 0.f; acc1[b] = 0.f; }
    const int nch = K >> 3;
    int c = lane;
    for (; c + 96 < nch; c += 128) {
        uint4 a0 = ldg_nc_v4(r0 + (size_t)c * 8), a1 = ldg_nc_v4(r0 + (size_t)(c + 32) * 8);
        uint4 a2 = ldg_nc_v4(r0 + (size_t)(c + 64) * 8), a3 = ldg_nc_v4(r0 + (size_t)(c + 96) * 8);
        uint4 b0 = ldg_nc_v4(r1 + (size_t)c * 8), b1 = ldg_nc_v4(r1 + (size_t)(c + 32) * 8);
        uint4 b2 = ldg_nc_v4(r1 + (size_t)(c + 64) * 8), b3 = ldg_nc_v4(r1 + (size_t)(c + 96) * 8);
        dot8<NB>(a0, xs, K, c, acc0); dot8<NB>(a1, xs, K, c + 32, acc0);
        dot8<NB>(a2, xs, K, c + 64, acc0); dot8<NB>(a3, xs, K, c + 96, acc0);
        dot8<NB>(b0, xs, K, c, acc1); dot8<NB>(b1, xs, K, c + 32, acc1);
        dot8<NB>(b2, xs, K, c + 64, acc1); dot8<NB>(b3, xs, K, c + 96, acc1);
    }
    for (; c < nch; c += 32) {
        uint4 a0 = ldg_nc_v4(r0 + (size_t)c * 8);
        uint4 b0 = ldg_nc_v4(r1 + (size_t)c * 8);
        dot8<NB>(a0, xs, K, c, acc0);
        dot8<NB>(b0, xs, K, c, acc1);
    }
#pragma unroll
    for (int b = 0; b < NB; ++b) { acc0[b] = warp_sum(acc0[b]); acc1[b] = warp_sum(acc1[b]); }
}

template <int NB, int PRO, int EPI>
__global__ void __launch_bounds__(GEMV_THREADS) gemv_kernel(GemvArgs a) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    bf16* xs = reinterpret_cast<bf16*>(smem_raw);  // [NB][K]
    __shared__ float s_val[GEMV_WARPS][NB];
    __shared__ int s_idx[GEMV_WARPS][NB];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int K = a.K;

    // ---- prologue: stage the activation rows
    if (PRO == PRO_RMSNORM) {
        for (int b = 0; b < NB; ++b) {
            if (b < a.B) {
                const bf16* xr = a.x + (size_t)b * K;
                float rstd = rms_rstd_warp(xr, K, a.eps, lane);      // every warp computes the same value (same order)
                for (int c = tid * 8; c < K; c += GEMV_THREADS * 8) {
                    uint4 o = rms_apply8(*reinterpret_cast<const uint4*>(xr + c), *reinterpret_cast<const uint4*>(a.ln_w + c), rstd);
                    *reinterpret_cast<uint4*>(xs + (size_t)b * K + c) = o;
                    if (EPI == GE_LOGITS && blockIdx.x == 0) {
                        *reinterpret_cast<uint4*>(a.last_hidden + (size_t)b * K + c) = o;
                        if (a.hidden_steps) {
                            int t = a.state[0];   // tokens sampled so far == index of the hidden state in the log
                            if (t < a.max_steps) *reinterpret_cast<uint4*>(a.hidden_steps + ((size_t)t * a.B + b) * K + c) = o;
                        }
                    }
                }
            } else {
                for (int c = tid * 8; c < K; c += GEMV_THREADS * 8) *reinterpret_cast<uint4*>(xs + (size_t)b * K + c) = make_uint4(0, 0, 0, 0);
            }
        }
    } else {
        for (int b = 0; b < NB; ++b)
            for (int c = tid * 8; c < K; c += GEMV_THREADS * 8)
                *reinterpret_cast<uint4*>(xs + (size_t)b * K + c) =
                    b < a.B ? *reinterpret_cast<const uint4*>(a.x + (size_t)b * K + c) : make_uint4(0, 0, 0, 0);
    }
    __syncthreads();

    float best_v[NB];
    int best_i[NB];
    int vlo[NB], vhi[NB];
    if (EPI == GE_LOGITS) {
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            best_v[b] = -INFINITY; best_i[b] = 0x7fffffff;
            vlo[b] = b < a.B ? a.state[st_vlo(b)] : 0;
            vhi[b] = b < a.B ? a.state[st_vhi(b)] : 0;
        }
    }

    const int n_pairs = (a.N + 1) >> 1;
    for (int pair = blockIdx.x * GEMV_WARPS + warp; pair < n_pairs; pair += gridDim.x * GEMV_WARPS) {
        int n0 = pair * 2, n1 = n0 + 1;
        bool has1 = n1 < a.N;
        const bf16 *r0, *r1;
        if (EPI == GE_LOGITS) {
            r0 = n0 < a.vocab ? a.W + (size_t)n0 * K : a.W2 + (size_t)(n0 - a.vocab) * K;
            int nn = has1 ? n1 : n0;
            r1 = nn < a.vocab ? a.W + (size_t)nn * K : a.W2 + (size_t)(nn - a.vocab) * K;
        } else {
            r0 = a.W + (size_t)n0 * K;
            r1 = a.W + (size_t)(has1 ? n1 : n0) * K;
        }
        if (EPI == GE_LOGITS && n0 >= a.vocab) {
            // prototype rows no batch row may emit cannot win the arg-max: do not stream them (multi-image tables)
            bool any = false;
#pragma unroll
            for (int b = 0; b < NB; ++b) any |= (n0 - a.vocab + 1 >= vlo[b] && n0 - a.vocab < vhi[b]);
            if (!any) continue;
        }
        float acc0[NB], acc1[NB];
        warp_rows2<NB>(r0, r1, K, xs, lane, acc0, acc1);
        if (EPI == GE_LOGITS) {
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                // bf16-rounded logit (padt.py:294), -inf outside the row's own VRT range (padt.py:301), first max wins
                float l0 = rbf(acc0[b]), l1 = rbf(acc1[b]);
                bool ok0 = n0 < a.vocab || (n0 - a.vocab >= vlo[b] && n0 - a.vocab < vhi[b]);
                bool ok1 = has1 && (n1 < a.vocab || (n1 - a.vocab >= vlo[b] && n1 - a.vocab < vhi[b]));
                if (ok0 && (l0 > best_v[b] || (l0 == best_v[b] && n0 < best_i[b]))) { best_v[b] = l0; best_i[b] = n0; }
                if (ok1 && (l1 > best_v[b] || (l1 == best_v[b] && n1 < best_i[b]))) { best_v[b] = l1; best_i[b] = n1; }
            }
        } else if (lane == 0) {
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                if (b >= a.B) break;
                if (EPI == GE_BIAS) {
                    float v0 = acc0[b] + (a.bias ? bf2f(a.bias[n0]) : 0.f);
                    a.y[(size_t)b * a.N + n0] = f2bf(v0);
                    if (has1) a.y[(size_t)b * a.N + n1] = f2bf(acc1[b] + (a.bias ? bf2f(a.bias[n1]) : 0.f));
                } else if (EPI == GE_RESID) {
                    bf16* h0 = a.y + (size_t)b * a.N + n0;
                    h0[0] = f2bf(bf2f(h0[0]) + rbf(acc0[b]));
                    if (has1) h0[1] = f2bf(bf2f(h0[1]) + rbf(acc1[b]));
                } else if (EPI == GE_SWIGLU) {
                    float g = rbf(acc0[b]), u = rbf(acc1[b]);
                    a.y[(size_t)b * (a.N >> 1) + pair] = f2bf(rbf(silu_f(g)) * u);
                }
            }
        }
    }

    if (EPI == GE_LOGITS) {
        if (lane == 0) {
#pragma unroll
            for (int b = 0; b < NB; ++b) { s_val[warp][b] = best_v[b]; s_idx[warp][b] = best_i[b]; }
        }
        __syncthreads();
        if (tid < a.B) {
            float bv = -INFINITY; int bi = 0x7fffffff;
            for (int w = 0; w < GEMV_WARPS; ++w) {
                float v = s_val[w][tid]; int i = s_idx[w][tid];
                if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
            }
            a.amax_val[(size_t)blockIdx.x * a.B + tid] = bv;
            a.amax_idx[(size_t)blockIdx.x * a.B + tid] = bi;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// sample + bookkeeping + next-token embedding gather   (padt.py:745-757 and padt.py:226-229)
// ---------------------------------------------------------------------------------------------------------------
struct SampleArgs {
    const float* amax_val; const int* amax_idx; int n_blocks;
    int B, D, vocab, n_protos;
    int* state;
    long long* cur_token; long long* tokens_out; long long* argmax_out; const long long* forced; int max_steps;
    const long long* eos_ids; int n_eos; long long pad_id;
    const bf16 *embed, *protos; bf16* h;
};

__global__ void __launch_bounds__(256) sample_embed_kernel(SampleArgs a) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    __shared__ long long s_tok[ST_BMAX];
    const int t = a.state[0];
    for (int b = warp; b < a.B; b += 8) {
        float bv = -INFINITY; int bi = 0x7fffffff;
        for (int k = lane; k < a.n_blocks; k += 32) {
            float v = a.amax_val[(size_t)k * a.B + b]; int i = a.amax_idx[(size_t)k * a.B + b];
            if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, bv, o); int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) {
            long long arg = bi;
            long long tok = arg;
            if (a.forced && t < a.max_steps) tok = a.forced[(size_t)t * a.B + b];
            int unf = a.state[st_unfin(b)];
            if (a.n_eos > 0) tok = unf ? tok : a.pad_id;
            bool is_eos = false;
            for (int e = 0; e < a.n_eos; ++e) is_eos |= (tok == a.eos_ids[e]);
            if (is_eos) unf = 0;
            a.state[st_unfin(b)] = unf;
            if (t < a.max_steps) {
                a.tokens_out[(size_t)t * a.B + b] = tok;
                if (a.argmax_out) a.argmax_out[(size_t)t * a.B + b] = arg;
            }
            a.cur_token[b] = tok;
            s_tok[b] = tok;
        }
    }
    __syncthreads();
    if (tid == 0) {
        int n = 0;
        for (int b = 0; b < a.B; ++b) n += a.state[st_unfin(b)];
        a.state[1] = n;
        a.state[0] = t + 1;
    }
    const int chunks = a.D >> 3;
    for (int i = tid; i < a.B * chunks; i += blockDim.x) {
        int b = i / chunks, c = i % chunks;
        long long id = s_tok[b];
        const bf16* src = (id < a.vocab) ? a.embed + (size_t)id * a.D : a.protos + (size_t)(id - a.vocab) * a.D;
        *reinterpret_cast<uint4*>(a.h + (size_t)b * a.D + c * 8) = *reinterpret_cast<const uint4*>(src + c * 8);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// decode attention: RoPE(q, k_new) -> KV append -> split-KV partial softmax -> last CTA of each (row, kv head) combines
// ---------------------------------------------------------------------------------------------------------------
constexpr int DA_THREADS = 256, DA_TILE = 64, DA_HD = 128, DA_LD = DA_HD + 8, DA_MAXG = 8;

struct DecAttnArgs {
    const bf16* qkv;          // [B,(nh+2nkv)*hd] raw projections of this step
    bf16 *kcache, *vcache;    // this layer: [B][nkv][s_max][hd]
    const bf16 *cos_tab, *sin_tab; int max_pos, sec0, sec1;
    const int* state;
    float *part_o, *part_ml; int* counters;
    bf16* out;                // [B, nh*hd]
    int B, nh, nkv, s_max, n_split;
    float scale_log2;
};

__global__ void __launch_bounds__(DA_THREADS) decode_attn_kernel(DecAttnArgs a) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    bf16* sK = reinterpret_cast<bf16*>(smem_raw);           // [64][136]
    bf16* sV = sK + DA_TILE * DA_LD;                        // [64][136]
    float* sQ = reinterpret_cast<float*>(sV + DA_TILE * DA_LD);  // [group][128]
    float* sP = sQ + DA_MAXG * DA_HD;                       // [group][64]
    __shared__ int s_last;
    const int split = blockIdx.x, kvh = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int group = a.nh / a.nkv;
    const int heads_total = a.nh + 2 * a.nkv;
    const int s = a.state[0] - 1;            // index of the token being fed
    const int L = a.state[2];
    const int cur = L + s;                   // cache slot of the new key
    const int ks = a.state[st_kvstart(b)];
    int pos = cur + a.state[st_delta(b)];    // padt.py:270-276: cache_position[0] + rope_deltas, same on all 3 planes
    pos = pos < 0 ? 0 : (pos >= a.max_pos ? a.max_pos - 1 : pos);
    const int n_keys = cur + 1 - ks;
    int chunk = (n_keys + a.n_split - 1) / a.n_split;
    chunk = (chunk + 31) & ~31;
    const int k_begin = ks + split * chunk;
    int k_end = k_begin + chunk;
    k_end = k_end > cur + 1 ? cur + 1 : k_end;
    const bf16* row = a.qkv + (size_t)b * heads_total * DA_HD;
    bf16* kc = a.kcache + ((size_t)b * a.nkv + kvh) * a.s_max * DA_HD;
    bf16* vc = a.vcache + ((size_t)b * a.nkv + kvh) * a.s_max * DA_HD;

    // ---- roped queries of this kv group -> sQ (fp32 values of the bf16-rounded result)
    for (int i = tid; i < group * 64; i += DA_THREADS) {
        int g = i >> 6, j = i & 63;
        const bf16* q = row + (size_t)(kvh * group + g) * DA_HD;
        float c = bf2f(a.cos_tab[(size_t)pos * 64 + j]), sn = bf2f(a.sin_tab[(size_t)pos * 64 + j]);
        float lo = bf2f(q[j]), hi = bf2f(q[j + 64]);
        sQ[g * DA_HD + j] = rbf(rbf(lo * c) + rbf(-hi * sn));
        sQ[g * DA_HD + j + 64] = rbf(rbf(hi * c) + rbf(lo * sn));
    }
    const bool owns_new = (cur >= k_begin && cur < k_end);

    float m_run = -INFINITY, l_run = 0.f;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int t0 = k_begin; t0 < k_end; t0 += DA_TILE) {
        __syncthreads();   // previous tile fully consumed (also orders the sQ writes on the first pass)
        for (int id = tid; id < DA_TILE * 16; id += DA_THREADS) {
            int r = id >> 4, c = id & 15;
            int key = t0 + r;
            bool ok = key < k_end && key != cur;
            cp_async16(sK + r * DA_LD + c * 8, kc + (size_t)(ok ? key : 0) * DA_HD + c * 8, ok ? 16 : 0);
            cp_async16(sV + r * DA_LD + c * 8, vc + (size_t)(ok ? key : 0) * DA_HD + c * 8, ok ? 16 : 0);
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        if (owns_new && cur >= t0 && cur < t0 + DA_TILE) {
            // new key/value of this step: rope k in bf16 semantics, append to the cache, and use it from smem
            int r = cur - t0;
            const bf16* kn = row + (size_t)(a.nh + kvh) * DA_HD;
            const bf16* vn = row + (size_t)(a.nh + a.nkv + kvh) * DA_HD;
            if (tid < 64) {
                int j = tid;
                float c = bf2f(a.cos_tab[(size_t)pos * 64 + j]), sn = bf2f(a.sin_tab[(size_t)pos * 64 + j]);
                float lo = bf2f(kn[j]), hi = bf2f(kn[j + 64]);
                bf16 olo = f2bf(rbf(lo * c) + rbf(-hi * sn)), ohi = f2bf(rbf(hi * c) + rbf(lo * sn));
                sK[r * DA_LD + j] = olo; sK[r * DA_LD + j + 64] = ohi;
                kc[(size_t)cur * DA_HD + j] = olo; kc[(size_t)cur * DA_HD + j + 64] = ohi;
            } else if (tid < 192) {
                int j = tid - 64;
                bf16 v = vn[j];
                sV[r * DA_LD + j] = v;
                vc[(size_t)cur * DA_HD + j] = v;
            }
            __syncthreads();
        }
        if (warp < group) {
            const float* q = sQ + warp * DA_HD;
            float sc[2];
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                int r = lane + kk * 32;
                const bf16* kr = sK + r * DA_LD;
                float d = 0.f;
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    uint4 kv = *reinterpret_cast<const uint4*>(kr + c * 8);
                    float4 q0 = *reinterpret_cast<const float4*>(q + c * 8), q1 = *reinterpret_cast<const float4*>(q + c * 8 + 4);
                    float2 k0 = bf2x_to_f2(kv.x), k1 = bf2x_to_f2(kv.y), k2 = bf2x_to_f2(kv.z), k3 = bf2x_to_f2(kv.w);
                    d = fmaf(k0.x, q0.x, d); d = fmaf(k0.y, q0.y, d); d = fmaf(k1.x, q0.z, d); d = fmaf(k1.y, q0.w, d);
                    d = fmaf(k2.x, q1.x, d); d = fmaf(k2.y, q1.y, d); d = fmaf(k3.x, q1.z, d); d = fmaf(k3.y, q1.w, d);
                }
                sc[kk] = (t0 + r < k_end) ? d * a.scale_log2 : -INFINITY;
            }
            float mx = warp_max(fmaxf(sc[0], sc[1]));
            float m_new = fmaxf(m_run, mx);
            float m_use = (m_new == -INFINITY) ? 0.f : m_new;
            float alpha = exp2f(m_run - m_use);
            float p0 = exp2f(sc[0] - m_use), p1 = exp2f(sc[1] - m_use);
            l_run = l_run * alpha + warp_sum(p0 + p1);
            m_run = m_new;
            sP[warp * DA_TILE + lane] = rbf(p0);          // FA2: P rounded to bf16 before P@V
            sP[warp * DA_TILE + lane + 32] = rbf(p1);
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[k] *= alpha;
            const float* pw = sP + warp * DA_TILE;
            for (int r = 0; r < DA_TILE; ++r) {
                float p = pw[r];
                uint2 vv = *reinterpret_cast<const uint2*>(sV + r * DA_LD + lane * 4);
                float2 v0 = bf2x_to_f2(vv.x), v1 = bf2x_to_f2(vv.y);
                acc[0] = fmaf(p, v0.x, acc[0]); acc[1] = fmaf(p, v0.y, acc[1]);
                acc[2] = fmaf(p, v1.x, acc[2]); acc[3] = fmaf(p, v1.y, acc[3]);
            }
            __syncwarp();
        }
    }
    // ---- write the partial of this split
    if (warp < group) {
        int head = kvh * group + warp;
        size_t base = ((size_t)b * a.nh + head) * a.n_split + split;
        *reinterpret_cast<float4*>(a.part_o + base * DA_HD + lane * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        if (lane == 0) { a.part_ml[base * 2] = m_run; a.part_ml[base * 2 + 1] = l_run; }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        int prev = atomicAdd(a.counters + b * a.nkv + kvh, 1);
        s_last = (prev == a.n_split - 1);
        if (s_last) a.counters[b * a.nkv + kvh] = 0;   // re-arm for the next launch
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (warp < group) {
        int head = kvh * group + warp;
        size_t base = ((size_t)b * a.nh + head) * a.n_split;
        float M = -INFINITY;
        for (int c = 0; c < a.n_split; ++c) M = fmaxf(M, __ldcg(a.part_ml + (base + c) * 2));
        float Ls = 0.f, o[4] = {0.f, 0.f, 0.f, 0.f};
        for (int c = 0; c < a.n_split; ++c) {
            float m = __ldcg(a.part_ml + (base + c) * 2), l = __ldcg(a.part_ml + (base + c) * 2 + 1);
            if (m == -INFINITY) continue;
            float w = exp2f(m - M);
            Ls += l * w;
            float4 po = __ldcg(reinterpret_cast<const float4*>(a.part_o + (base + c) * DA_HD + lane * 4));
            o[0] += po.x * w; o[1] += po.y * w; o[2] += po.z * w; o[3] += po.w * w;
        }
        float inv = Ls > 0.f ? 1.0f / Ls : 0.f;
        uint2 pk;
        pk.x = f2_to_bf2x(o[0] * inv, o[1] * inv);
        pk.y = f2_to_bf2x(o[2] * inv, o[3] * inv);
        *reinterpret_cast<uint2*>(a.out + (size_t)b * a.nh * DA_HD + (size_t)head * DA_HD + lane * 4) = pk;
    }
}

template <int NB, int PRO, int EPI>
int launch_gemv_t(const GemvArgs& a, int grid, cudaStream_t st) {
    size_t smem = (size_t)NB * a.K * sizeof(bf16);
    static size_t configured = 0;
    if (smem > configured) {
        PADT_CHECK_CUDA(cudaFuncSetAttribute(gemv_kernel<NB, PRO, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem > 48 * 1024 ? smem : 48 * 1024)));
        configured = smem > 48 * 1024 ? smem : 48 * 1024;
    }
    gemv_kernel<NB, PRO, EPI><<<grid, GEMV_THREADS, smem, st>>>(a);
    PADT_LAUNCH_CHECK();
    return 0;
}

template <int PRO, int EPI>
int launch_gemv(const GemvArgs& a, int grid, cudaStream_t st) {
    PADT_REQUIRE(a.K % 8 == 0, "K must be a multiple of 8");
    PADT_REQUIRE((size_t)a.B * a.K * 2 <= 200 * 1024, "activation rows do not fit in shared memory");
    if (a.B == 1) return launch_gemv_t<1, PRO, EPI>(a, grid, st);
    if (a.B == 2) return launch_gemv_t<2, PRO, EPI>(a, grid, st);
    if (a.B <= 4) return launch_gemv_t<4, PRO, EPI>(a, grid, st);
    if (a.B <= 8) return launch_gemv_t<8, PRO, EPI>(a, grid, st);
    padt_set_error("decode batch > 8 rows per GPU is not supported by the v1 GEMV path");
    return -1;
}

inline int gemv_grid(int N) {
    int pairs = (N + 1) / 2;
    int g = cdiv(pairs, GEMV_WARPS);
    int cap = 148 * 8;
    return g < cap ? g : cap;
}

}  // namespace

int decode_logits_blocks(int n_rows) { return gemv_grid(n_rows); }

int launch_sample_embed(const DecodeParams& p, cudaStream_t st) {
    SampleArgs a;
    a.amax_val = p.amax_val; a.amax_idx = p.amax_idx; a.n_blocks = p.n_amax_blocks;
    a.B = p.B; a.D = p.D; a.vocab = p.vocab; a.n_protos = p.n_protos;
    a.state = p.dev_state; a.cur_token = p.cur_token; a.tokens_out = p.tokens_out; a.argmax_out = p.argmax_out;
    a.forced = p.forced; a.max_steps = p.max_steps;
    a.eos_ids = p.eos_ids; a.n_eos = p.n_eos; a.pad_id = p.pad_id;
    a.embed = p.embed; a.protos = p.protos; a.h = p.h;
    sample_embed_kernel<<<1, 256, 0, st>>>(a);
    PADT_LAUNCH_CHECK();
    return 0;
}

// final RMSNorm (HF:937) + VRT-extended logits + masked arg-max partials; `h` holds the pre-norm rows [B,D]
int launch_final_logits(const DecodeParams& p, cudaStream_t st) {
    GemvArgs a = {};
    a.x = p.h; a.ln_w = p.final_norm; a.eps = p.eps;
    a.W = p.head; a.W2 = p.protos; a.B = p.B; a.N = p.vocab + p.n_protos; a.K = p.D; a.vocab = p.vocab;
    a.state = p.dev_state; a.amax_val = p.amax_val; a.amax_idx = p.amax_idx;
    a.last_hidden = p.last_hidden; a.hidden_steps = p.hidden_steps; a.max_steps = p.max_steps;
    PADT_REQUIRE(p.n_amax_blocks == gemv_grid(a.N), "arg-max partial buffer size mismatch");
    return launch_gemv<PRO_RMSNORM, GE_LOGITS>(a, p.n_amax_blocks, st);
}

int launch_decode_layers(const DecodeParams& p, cudaStream_t st) {
    PADT_REQUIRE(p.hd == DA_HD, "decode attention is specialised for head_dim 128");
    PADT_REQUIRE(p.nh % p.nkv == 0 && p.nh / p.nkv <= DA_MAXG, "GQA group must be <= 8");
    PADT_REQUIRE(p.B <= ST_BMAX, "batch");
    const int qkv_n = (p.nh + 2 * p.nkv) * p.hd;
    const size_t layer_kv = (size_t)p.B * p.nkv * p.s_max * p.hd;
    static bool attn_configured = false;
    const int attn_smem = 2 * DA_TILE * DA_LD * 2 + DA_MAXG * DA_HD * 4 + DA_MAXG * DA_TILE * 4;
    if (!attn_configured) {
        PADT_CHECK_CUDA(cudaFuncSetAttribute(decode_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_smem));
        attn_configured = true;
    }
    for (int li = 0; li < p.n_layers; ++li) {
        const DecodeLayerW& w = p.layers[li];
        {   // RMSNorm -> fused q/k/v projection + bias (HF:807, 724-726)
            GemvArgs a = {};
            a.x = p.h; a.ln_w = w.ln1; a.eps = p.eps; a.W = w.wqkv; a.bias = w.bqkv; a.y = p.qkv;
            a.B = p.B; a.N = qkv_n; a.K = p.D;
            int rc = launch_gemv<PRO_RMSNORM, GE_BIAS>(a, gemv_grid(a.N), st);
            if (rc) return rc;
        }
        {   // M-RoPE + cache append + attention (HF:732-755)
            DecAttnArgs a;
            a.qkv = p.qkv; a.kcache = p.kcache + li * layer_kv; a.vcache = p.vcache + li * layer_kv;
            a.cos_tab = p.cos_tab; a.sin_tab = p.sin_tab; a.max_pos = p.max_pos; a.sec0 = 0; a.sec1 = 0;
            a.state = p.dev_state; a.part_o = p.part_o; a.part_ml = p.part_ml; a.counters = p.attn_counters;
            a.out = p.act; a.B = p.B; a.nh = p.nh; a.nkv = p.nkv; a.s_max = p.s_max; a.n_split = p.n_split;
            a.scale_log2 = (1.0f / sqrtf((float)p.hd)) * 1.4426950408889634f;
            dim3 grid(p.n_split, p.nkv, p.B);
            decode_attn_kernel<<<grid, DA_THREADS, attn_smem, st>>>(a);
            PADT_LAUNCH_CHECK();
        }
        {   // o_proj + residual (HF:758, 819)
            GemvArgs a = {};
            a.x = p.act; a.W = w.wo; a.y = p.h; a.B = p.B; a.N = p.D; a.K = p.nh * p.hd;
            int rc = launch_gemv<PRO_NONE, GE_RESID>(a, gemv_grid(a.N), st);
            if (rc) return rc;
        }
        {   // RMSNorm -> interleaved gate/up -> SiLU*up (HF:823, 623)
            GemvArgs a = {};
            a.x = p.h; a.ln_w = w.ln2; a.eps = p.eps; a.W = w.wgu; a.y = p.act; a.B = p.B; a.N = 2 * p.I; a.K = p.D;
            int rc = launch_gemv<PRO_RMSNORM, GE_SWIGLU>(a, gemv_grid(a.N), st);
            if (rc) return rc;
        }
        {   // down_proj + residual (HF:623, 825)
            GemvArgs a = {};
            a.x = p.act; a.W = w.wdown; a.y = p.h; a.B = p.B; a.N = p.D; a.K = p.I;
            int rc = launch_gemv<PRO_NONE, GE_RESID>(a, gemv_grid(a.N), st);
            if (rc) return rc;
        }
    }
    return 0;
}
