// Device helpers shared by the per-kernel decode path (decode.cu) and the persistent decode kernel (decode_mega.cu).
// Both paths use the same chunk order / FMA order / butterfly reductions, so they produce bit-identical projections.
#pragma once
#include "common.cuh"
#include "kernels.h"

namespace {

// state block layout: see kernels.h
constexpr int ST_HDR = PADT_STATE_HDR, ST_BMAX = PADT_STATE_BMAX;
__device__ __host__ inline int st_kvstart(int b) { return ST_HDR + 0 * ST_BMAX + b; }
__device__ __host__ inline int st_delta(int b) { return ST_HDR + 1 * ST_BMAX + b; }
__device__ __host__ inline int st_vlo(int b) { return ST_HDR + 2 * ST_BMAX + b; }
__device__ __host__ inline int st_vhi(int b) { return ST_HDR + 3 * ST_BMAX + b; }
__device__ __host__ inline int st_unfin(int b) { return ST_HDR + 4 * ST_BMAX + b; }

// 8 weights x NB activation rows.  Two independent accumulators per output (even / odd elements) halve the dependent FMA
// chain; every decode path goes through this function, so they all share one summation order.
template <int NB>
__device__ __forceinline__ void dot8(const uint4& w, const bf16* xs, int K, int c, float2* acc) {
    float2 w0 = bf2x_to_f2(w.x), w1 = bf2x_to_f2(w.y), w2 = bf2x_to_f2(w.z), w3 = bf2x_to_f2(w.w);
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        uint4 xv = *reinterpret_cast<const uint4*>(xs + (size_t)b * K + c * 8);
        float2 x0 = bf2x_to_f2(xv.x), x1 = bf2x_to_f2(xv.y), x2 = bf2x_to_f2(xv.z), x3 = bf2x_to_f2(xv.w);
        float2 a = acc[b];
        a.x = fmaf(w0.x, x0.x, a.x); a.y = fmaf(w0.y, x0.y, a.y);
        a.x = fmaf(w1.x, x1.x, a.x); a.y = fmaf(w1.y, x1.y, a.y);
        a.x = fmaf(w2.x, x2.x, a.x); a.y = fmaf(w2.y, x2.y, a.y);
        a.x = fmaf(w3.x, x3.x, a.x); a.y = fmaf(w3.y, x3.y, a.y);
        acc[b] = a;
    }
}

// One warp computes two weight rows (global memory) against NB activation rows held in shared memory.
template <int NB>
__device__ __forceinline__ void warp_rows2(const bf16* __restrict__ r0, const bf16* __restrict__ r1, int K, const bf16* xs, int lane,
                                           float* out0, float* out1) {
    float2 acc0[NB], acc1[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) { acc0[b] = make_float2(0.f, 0.f); acc1[b] = make_float2(0.f, 0.f); }
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
    for (int b = 0; b < NB; ++b) { out0[b] = warp_sum(acc0[b].x + acc0[b].y); out1[b] = warp_sum(acc1[b].x + acc1[b].y); }
}

struct BlockSync { __device__ __forceinline__ void operator()() const { __syncthreads(); } };
// named barrier over the first `N` threads of the block (the persistent kernel keeps its producer warp out of it)
template <int ID, int N>
struct NamedSync { __device__ __forceinline__ void operator()() const { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(N) : "memory"); } };

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

template <class Sync>
__device__ __forceinline__ void sample_embed_body(const SampleArgs& a, int tid, int nthreads, Sync sync) {
    const int warp = tid >> 5, lane = tid & 31;
    __shared__ long long s_tok[ST_BMAX];
    const int t = a.state[0];
    for (int b = warp; b < a.B; b += 8) {
        float bv = -INFINITY; int bi = 0x7fffffff;
        const int n_part = a.state[3];   // written by whichever kernel produced the arg-max partials
        for (int k = lane; k < n_part; k += 32) {
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
    sync();
    if (tid == 0) {
        int n = 0;
        for (int b = 0; b < a.B; ++b) n += a.state[st_unfin(b)];
        a.state[1] = n;
        a.state[0] = t + 1;
    }
    const int chunks = a.D >> 3;
    for (int i = tid; i < a.B * chunks; i += nthreads) {
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

// key range [k_begin, k_end) of one (row, split) item; `cur` = cache slot of the key appended in this step
__device__ __forceinline__ void decode_attn_range(const DecAttnArgs& a, int split, int b, int& k_begin, int& k_end, int& cur) {
    const int s = a.state[0] - 1;            // index of the token being fed
    cur = a.state[2] + s;
    const int ks = a.state[st_kvstart(b)];
    const int n_keys = cur + 1 - ks;
    int chunk = (n_keys + a.n_split - 1) / a.n_split;
    chunk = (chunk + 31) & ~31;
    k_begin = ks + split * chunk;
    k_end = k_begin + chunk;
    k_end = k_end > cur + 1 ? cur + 1 : k_end;
}

// cp.async one K/V tile (keys t0 .. t0+63 of the item) into the padded shared-memory tiles; the key of the current step is
// not in the cache yet and is zero-filled (its owner writes it from the fresh projection).  Commits one cp.async group.
__device__ __forceinline__ void decode_attn_issue_tile(const DecAttnArgs& a, int kvh, int b, int t0, int k_end, int cur, uint8_t* smem_raw, int tid,
                                                       int nthreads) {
    bf16* sK = reinterpret_cast<bf16*>(smem_raw);
    bf16* sV = sK + DA_TILE * DA_LD;
    const bf16* kc = a.kcache + ((size_t)b * a.nkv + kvh) * a.s_max * DA_HD;
    const bf16* vc = a.vcache + ((size_t)b * a.nkv + kvh) * a.s_max * DA_HD;
    for (int id = tid; id < DA_TILE * 16; id += nthreads) {
        int r = id >> 4, c = id & 15;
        int key = t0 + r;
        bool ok = key < k_end && key != cur;
        cp_async16(sK + r * DA_LD + c * 8, kc + (size_t)(ok ? key : 0) * DA_HD + c * 8, ok ? 16 : 0);
        cp_async16(sV + r * DA_LD + c * 8, vc + (size_t)(ok ? key : 0) * DA_HD + c * 8, ok ? 16 : 0);
    }
    cp_async_commit();
}

// `first_tile_in_flight`: the caller already issued decode_attn_issue_tile for the item's first tile (the persistent kernel
// does so before the grid barrier that publishes q/k/v, so the cache rows are in shared memory when the barrier opens).
// COMBINE = true: the last-arriving CTA of a (row, kv head) combines all split partials (per-layer kernel path);
// COMBINE = false: return after the partial (m, l, acc) of this split is written -- the caller combines (persistent kernel).
template <bool COMBINE, class Sync>
__device__ __forceinline__ void decode_attn_body(const DecAttnArgs& a, int split, int kvh, int b, uint8_t* smem_raw, int tid, int nthreads,
                                                 Sync sync, bool first_tile_in_flight = false) {
    bf16* sK = reinterpret_cast<bf16*>(smem_raw);           // [64][136]
    bf16* sV = sK + DA_TILE * DA_LD;                        // [64][136]
    float* sQ = reinterpret_cast<float*>(sV + DA_TILE * DA_LD);  // [group][128]
    float* sP = sQ + DA_MAXG * DA_HD;                       // [group][64]
    __shared__ int s_last;
    const int warp = tid >> 5, lane = tid & 31;
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
    for (int i = tid; i < group * 64; i += nthreads) {
        int g = i >> 6, j = i & 63;
        const bf16* q = row + (size_t)(kvh * group + g) * DA_HD;
        float c = bf2f(a.cos_tab[(size_t)pos * 64 + j]), sn = bf2f(a.sin_tab[(size_t)pos * 64 + j]);
        float lo = bf2f(q[j]), hi = bf2f(q[j + 64]);
        sQ[g * DA_HD + j] = rbf(rbf(lo * c) + rbf(-hi * sn));
        sQ[g * DA_HD + j + 64] = rbf(rbf(hi * c) + rbf(lo * sn));
    }
    const bool owns_new = (cur >= k_begin && cur < k_end);
    // the split that owns the key of this step loads its raw k / v (and rotary factors) together with the queries, so that
    // the append below does not add a dependent global round trip after the K/V tile wait (every other split of the kv head
    // waits for this one before the partials can be combined)
    const bf16* kn = row + (size_t)(a.nh + kvh) * DA_HD;
    const bf16* vn = row + (size_t)(a.nh + a.nkv + kvh) * DA_HD;
    float nk_lo = 0.f, nk_hi = 0.f, nk_c = 0.f, nk_s = 0.f;
    bf16 nv = f2bf(0.f);
    if (owns_new) {
        if (tid < 64) {
            nk_c = bf2f(a.cos_tab[(size_t)pos * 64 + tid]); nk_s = bf2f(a.sin_tab[(size_t)pos * 64 + tid]);
            nk_lo = bf2f(kn[tid]); nk_hi = bf2f(kn[tid + 64]);
        } else if (tid < 192) {
            nv = vn[tid - 64];
        }
    }

    float m_run = -INFINITY, l_run = 0.f;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int t0 = k_begin; t0 < k_end; t0 += DA_TILE) {
        sync();   // previous tile fully consumed (also orders the sQ writes on the first pass)
        if (!(first_tile_in_flight && t0 == k_begin)) decode_attn_issue_tile(a, kvh, b, t0, k_end, cur, smem_raw, tid, nthreads);
        cp_async_wait<0>();
        sync();
        if (owns_new && cur >= t0 && cur < t0 + DA_TILE) {
            // new key/value of this step: rope k in bf16 semantics, append to the cache, and use it from smem
            int r = cur - t0;
            if (tid < 64) {
                int j = tid;
                bf16 olo = f2bf(rbf(nk_lo * nk_c) + rbf(-nk_hi * nk_s)), ohi = f2bf(rbf(nk_hi * nk_c) + rbf(nk_lo * nk_s));
                sK[r * DA_LD + j] = olo; sK[r * DA_LD + j + 64] = ohi;
                kc[(size_t)cur * DA_HD + j] = olo; kc[(size_t)cur * DA_HD + j + 64] = ohi;
            } else if (tid < 192) {
                int j = tid - 64;
                sV[r * DA_LD + j] = nv;
                vc[(size_t)cur * DA_HD + j] = nv;
            }
            sync();
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
    if (!COMBINE) return;
    __threadfence();
    sync();
    if (tid == 0) {
        int prev = atomicAdd(a.counters + b * a.nkv + kvh, 1);
        s_last = (prev == a.n_split - 1);
        if (s_last) a.counters[b * a.nkv + kvh] = 0;   // re-arm for the next launch
    }
    sync();
    if (s_last) {
        // ---- combine the split partials of this (row, kv head); with >= 2*group warps each head is shared by two warps
        __threadfence();
        float (*s_comb)[32][6] = reinterpret_cast<float (*)[32][6]>(smem_raw);   // K/V tiles are dead by now: reuse them
        const int nw = nthreads >> 5;
        const int halves = nw >= 2 * group ? 2 : 1;
        const int hl = warp % group, half = warp / group;
        const bool active = warp < group * halves;
        const int per = (a.n_split + halves - 1) / halves;
        const int c_lo = half * per, c_hi = (c_lo + per) < a.n_split ? (c_lo + per) : a.n_split;
        const int head = kvh * group + hl;
        const size_t base = ((size_t)b * a.nh + head) * a.n_split;
        float M = -INFINITY, Ls = 0.f, o[4] = {0.f, 0.f, 0.f, 0.f};
        if (active) {
            const float* ml = a.part_ml + base * 2;
            int c0 = c_lo + lane, c1 = c_lo + lane + 32;
            float m0 = c0 < c_hi ? __ldcg(ml + c0 * 2) : -INFINITY, l0 = c0 < c_hi ? __ldcg(ml + c0 * 2 + 1) : 0.f;
            float m1 = c1 < c_hi ? __ldcg(ml + c1 * 2) : -INFINITY, l1 = c1 < c_hi ? __ldcg(ml + c1 * 2 + 1) : 0.f;
            M = warp_max(fmaxf(m0, m1));
            float w0 = (m0 == -INFINITY) ? 0.f : exp2f(m0 - M), w1 = (m1 == -INFINITY) ? 0.f : exp2f(m1 - M);
            Ls = warp_sum(l0 * w0 + l1 * w1);
            const float4* po = reinterpret_cast<const float4*>(a.part_o + base * DA_HD) + lane;
            for (int cb = c_lo; cb < c_hi; cb += 16) {
                float4 v[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) v[k] = (cb + k < c_hi) ? __ldcg(po + (size_t)(cb + k) * (DA_HD / 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    int c = cb + k - c_lo;
                    float w = __shfl_sync(0xffffffffu, c < 32 ? w0 : w1, c & 31);
                    o[0] += v[k].x * w; o[1] += v[k].y * w; o[2] += v[k].z * w; o[3] += v[k].w * w;
                }
            }
            if (half == 1) {
                float* d = s_comb[hl][lane];
                d[0] = o[0]; d[1] = o[1]; d[2] = o[2]; d[3] = o[3]; d[4] = M; d[5] = Ls;
            }
        }
        if (halves == 2) sync();
        if (active && half == 0) {
            if (halves == 2) {
                const float* d = s_comb[hl][lane];
                float M1 = d[4], L1 = d[5];
                float Mn = fmaxf(M, M1);
                float f0 = (M == -INFINITY) ? 0.f : exp2f(M - Mn), f1 = (M1 == -INFINITY) ? 0.f : exp2f(M1 - Mn);
                Ls = Ls * f0 + L1 * f1;
#pragma unroll
                for (int k = 0; k < 4; ++k) o[k] = o[k] * f0 + d[k] * f1;
            }
            float inv = Ls > 0.f ? 1.0f / Ls : 0.f;
            uint2 pk;
            pk.x = f2_to_bf2x(o[0] * inv, o[1] * inv);
            pk.y = f2_to_bf2x(o[2] * inv, o[3] * inv);
            *reinterpret_cast<uint2*>(a.out + (size_t)b * a.nh * DA_HD + (size_t)head * DA_HD + lane * 4) = pk;
        }
    }
}


}  // namespace
