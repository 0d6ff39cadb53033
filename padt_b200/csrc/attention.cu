// Segmented (varlen) flash attention, head_dim 80 / 128, optional causal + GQA + left-pad skip.
//
// Replaces flash_attn_varlen_func as the reference uses it: ViT window / full attention (HF:244-261 driven by
// padt.py:89-99, d=80), the PaDT decoder's 9 attention calls per vl_decode (padt_decoder.py:55, d=80, very
// asymmetric q/k lengths) and the LLM prefill attention (HF:744-755, causal GQA, d=128).
// Numerics follow FlashAttention-2: fp32 scores, online softmax in the exp2 domain, probabilities rounded to bf16
// before P@V, fp32 accumulation, one normalisation at the end.
//
// v1: mma.sync m16n8k16, 64-query x 64-key tiles, 4 warps, K/V double-buffered with cp.async.
#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int BQ = 64, BKV = 64, NTHREADS = 128;

template <int HD>
struct Cfg {
    static constexpr int LD = HD + 8;  // padded smem row (bf16) -> conflict-free ldmatrix
    static constexpr int TILE = BQ * LD;
    static constexpr int SMEM = (TILE + 4 * TILE) * 2;  // Q + 2x(K,V)
};

template <int HD>
__device__ __forceinline__ void load_tile(bf16* dst, const bf16* base, long long row_stride, int row0, int n_valid_rows, int tid) {
    // copies rows [row0, row0+64) x HD; rows >= n_valid_rows are zero filled
    constexpr int CH = HD / 8;
    constexpr int LD = Cfg<HD>::LD;
    for (int id = tid; id < BQ * CH; id += NTHREADS) {
        int r = id / CH, c = id % CH;
        int gr = row0 + r;
        bool ok = gr < n_valid_rows;
        const bf16* src = base + (long long)(ok ? gr : 0) * row_stride + c * 8;
        cp_async16(dst + r * LD + c * 8, src, ok ? 16 : 0);
    }
}

template <int HD, bool SPLIT>
__global__ void __launch_bounds__(NTHREADS) flash_attn_kernel(AttnParams p) {
    constexpr int LD = Cfg<HD>::LD;
    constexpr int TILE = Cfg<HD>::TILE;
    constexpr int KT = HD / 16;   // k-steps of Q K^T
    constexpr int DT = HD / 8;    // n8 tiles of the output
    extern __shared__ __align__(128) uint8_t smem_raw[];
    bf16* sQ = reinterpret_cast<bf16*>(smem_raw);
    bf16* sK = sQ + TILE;         // [2][TILE]
    bf16* sV = sK + 2 * TILE;     // [2][TILE]

    // causal: later query tiles attend over more keys -> schedule them first (CTAs are issued in blockIdx order), the short ones fill the tail
    const int split = SPLIT ? (int)blockIdx.x : 0;               // SPLIT: one query tile, blockIdx.x walks the key-range parts
    const int seg = blockIdx.z, h = blockIdx.y, qt = SPLIT ? 0 : (p.causal ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x);
    const int ql = p.q_len[seg];
    if (qt * BQ >= ql) return;
    const int kl = p.k_len[seg];
    const int ks = p.k_start ? p.k_start[seg] : 0;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int kvh = h / p.group;
    const bf16* qb = p.q + (long long)p.q_off[seg] * p.q_row + (long long)h * p.q_head;
    const bf16* kb = p.k + (long long)p.k_off[seg] * p.k_row + (long long)kvh * p.k_head;
    const bf16* vb = p.v + (long long)p.k_off[seg] * p.v_row + (long long)kvh * p.v_head;
    const int diag = kl - ql;  // causal: key j visible to query i iff j <= i + diag

    int kv_end = kl;
    if (p.causal) {
        int e = qt * BQ + BQ - 1 + diag + 1;
        kv_end = e < kl ? e : kl;
    }
    int t_begin = ks / BKV;
    int t_end = (kv_end + BKV - 1) / BKV;
    if (SPLIT) {
        const int per = (t_end - t_begin + p.n_split - 1) / p.n_split;
        t_begin += split * per;
        t_end = t_begin + per < t_end ? t_begin + per : t_end;
    }

    load_tile<HD>(sQ, qb, p.q_row, qt * BQ, ql, tid);
    if (t_begin < t_end) {
        load_tile<HD>(sK, kb, p.k_row, t_begin * BKV, kl, tid);
        load_tile<HD>(sV, vb, p.v_row, t_begin * BKV, kl, tid);
    }
    cp_async_commit();

    float o[DT][4];
#pragma unroll
    for (int i = 0; i < DT; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
    float m_run[2] = {-INFINITY, -INFINITY};
    float l_run[2] = {0.f, 0.f};
    const float sl2 = p.scale * 1.4426950408889634f;
    uint32_t qf[KT][4];

    for (int t = t_begin; t < t_end; ++t) {
        const int buf = (t - t_begin) & 1;
        // prefetch next K/V tile into the other buffer (its previous contents were consumed before the last barrier)
        if (t + 1 < t_end) {
            load_tile<HD>(sK + (buf ^ 1) * TILE, kb, p.k_row, (t + 1) * BKV, kl, tid);
            load_tile<HD>(sV + (buf ^ 1) * TILE, vb, p.v_row, (t + 1) * BKV, kl, tid);
        }
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        if (t == t_begin) {
#pragma unroll
            for (int kk = 0; kk < KT; ++kk) {
                int row = warp * 16 + (lane & 15);
                int col = kk * 16 + (lane >> 4) * 8;
                ldmatrix_x4(qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3], sQ + row * LD + col);
            }
        }
        const bf16* k_s = sK + buf * TILE;
        const bf16* v_s = sV + buf * TILE;
        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
        for (int kk = 0; kk < KT; ++kk) {
#pragma unroll
            for (int nj = 0; nj < 4; ++nj) {
                uint32_t b0, b1, b2, b3;
                int row = nj * 16 + (lane & 7) + ((lane >> 4) << 3);
                int col = kk * 16 + ((lane >> 3) & 1) * 8;
                ldmatrix_x4(b0, b1, b2, b3, k_s + row * LD + col);
                mma_bf16_16816(s[2 * nj], qf[kk], b0, b1);
                mma_bf16_16816(s[2 * nj + 1], qf[kk], b2, b3);
            }
        }
        // ---- scale, mask, online softmax
        const int i0 = qt * BQ + warp * 16 + g;
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                int j = t * BKV + ni * 8 + t4 * 2 + (r & 1);
                int i = i0 + (r >> 1) * 8;
                bool ok = (j < kl) && (j >= ks) && (!p.causal || j <= i + diag);
                float x = ok ? s[ni][r] * sl2 : -INFINITY;
                s[ni][r] = x;
                mx[r >> 1] = fmaxf(mx[r >> 1], x);
            }
        }
        float alpha[2], m_use[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            float v = mx[r];
            v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
            v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
            float m_new = fmaxf(m_run[r], v);
            m_use[r] = (m_new == -INFINITY) ? 0.f : m_new;
            alpha[r] = exp2f(m_run[r] - m_use[r]);
            m_run[r] = m_new;
        }
        float rs[2] = {0.f, 0.f};
        uint32_t pf[4][4];
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) {
            float p0 = exp2f(s[ni][0] - m_use[0]);
            float p1 = exp2f(s[ni][1] - m_use[0]);
            float p2 = exp2f(s[ni][2] - m_use[1]);
            float p3 = exp2f(s[ni][3] - m_use[1]);
            rs[0] += p0 + p1;
            rs[1] += p2 + p3;
            int kk2 = ni >> 1, hi = ni & 1;
            pf[kk2][hi * 2 + 0] = f2_to_bf2x(p0, p1);
            pf[kk2][hi * 2 + 1] = f2_to_bf2x(p2, p3);
        }
        l_run[0] = l_run[0] * alpha[0] + rs[0];
        l_run[1] = l_run[1] * alpha[1] + rs[1];
#pragma unroll
        for (int i = 0; i < DT; ++i) {
            o[i][0] *= alpha[0]; o[i][1] *= alpha[0];
            o[i][2] *= alpha[1]; o[i][3] *= alpha[1];
        }
        // ---- O += P V
#pragma unroll
        for (int kk2 = 0; kk2 < 4; ++kk2) {
#pragma unroll
            for (int dp = 0; dp < DT / 2; ++dp) {
                uint32_t b0, b1, b2, b3;
                int row = kk2 * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                int col = dp * 16 + (lane >> 4) * 8;
                ldmatrix_x4_trans(b0, b1, b2, b3, v_s + row * LD + col);
                mma_bf16_16816(o[2 * dp], pf[kk2], b0, b1);
                mma_bf16_16816(o[2 * dp + 1], pf[kk2], b2, b3);
            }
        }
        __syncthreads();  // all warps done with `buf` before the next iteration's prefetch overwrites it
    }
    cp_async_wait<0>();

    // ---- normalise + store
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        float l = l_run[r];
        l += __shfl_xor_sync(0xffffffffu, l, 1);
        l += __shfl_xor_sync(0xffffffffu, l, 2);
        l_run[r] = l;
    }
    if (SPLIT) {
        // un-normalised partial of this key-range part: rows of the single query tile
        const size_t base = (((size_t)seg * p.n_heads + h) * p.n_split + split) * BQ;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int i = warp * 16 + g + r * 8;
            float* po = p.part_o + (base + i) * HD;
#pragma unroll
            for (int dt = 0; dt < DT; ++dt) *reinterpret_cast<float2*>(po + dt * 8 + t4 * 2) = make_float2(o[dt][r * 2 + 0], o[dt][r * 2 + 1]);
            if (t4 == 0) { p.part_ml[(base + i) * 2] = m_run[r]; p.part_ml[(base + i) * 2 + 1] = l_run[r]; }
        }
        return;
    }
    bf16* ob = p.o + (long long)p.q_off[seg] * p.o_row + (long long)h * p.o_head;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        int i = qt * BQ + warp * 16 + g + r * 8;
        if (i >= ql) continue;
        float inv = l_run[r] > 0.f ? 1.0f / l_run[r] : 0.f;
        bf16* dst = ob + (long long)i * p.o_row;
#pragma unroll
        for (int dt = 0; dt < DT; ++dt) {
            float a = o[dt][r * 2 + 0] * inv, b = o[dt][r * 2 + 1] * inv;
            *reinterpret_cast<uint32_t*>(dst + dt * 8 + t4 * 2) = f2_to_bf2x(a, b);
        }
    }
}

// combine the key-range partials of one (segment, head): O = sum_s o_s 2^(m_s - M) / sum_s l_s 2^(m_s - M), M = max_s m_s (exp2 domain)
template <int HD>
__global__ void __launch_bounds__(128) attn_merge_kernel(AttnParams p) {
    const int seg = blockIdx.y, h = blockIdx.x;
    const int ql = p.q_len[seg];
    const size_t base = ((size_t)seg * p.n_heads + h) * p.n_split * BQ;
    bf16* ob = p.o + (long long)p.q_off[seg] * p.o_row + (long long)h * p.o_head;
    for (int idx = threadIdx.x; idx < ql * (HD / 2); idx += blockDim.x) {
        const int i = idx / (HD / 2), d = (idx % (HD / 2)) * 2;
        float M = -INFINITY;
        for (int s = 0; s < p.n_split; ++s) M = fmaxf(M, p.part_ml[(base + (size_t)s * BQ + i) * 2]);
        float L = 0.f, a = 0.f, b = 0.f;
        for (int s = 0; s < p.n_split; ++s) {
            const size_t row = base + (size_t)s * BQ + i;
            const float m = p.part_ml[row * 2];
            const float w = (m == -INFINITY) ? 0.f : exp2f(m - M);
            L += p.part_ml[row * 2 + 1] * w;
            const float2 v = *reinterpret_cast<const float2*>(p.part_o + row * HD + d);
            a += v.x * w; b += v.y * w;
        }
        const float inv = L > 0.f ? 1.0f / L : 0.f;
        *reinterpret_cast<uint32_t*>(ob + (long long)i * p.o_row + d) = f2_to_bf2x(a * inv, b * inv);
    }
}

template <int HD>
int launch_hd(const AttnParams& p, cudaStream_t st) {
    static bool configured_dev[64] = {false};
    bool& configured = configured_dev[padt_dev_slot()];
    if (!configured) {
        PADT_CHECK_CUDA(cudaFuncSetAttribute(flash_attn_kernel<HD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<HD>::SMEM));
        PADT_CHECK_CUDA(cudaFuncSetAttribute(flash_attn_kernel<HD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<HD>::SMEM));
        configured = true;
    }
    if (p.n_split > 1) {
        PADT_REQUIRE(p.max_q_len <= BQ && !p.causal && p.part_o && p.part_ml, "split-KV attention: one query tile, non-causal, workspaces");
        dim3 grid(p.n_split, p.n_heads, p.n_seg);
        flash_attn_kernel<HD, true><<<grid, NTHREADS, Cfg<HD>::SMEM, st>>>(p);
        PADT_LAUNCH_CHECK();
        attn_merge_kernel<HD><<<dim3(p.n_heads, p.n_seg), 128, 0, st>>>(p);
        PADT_LAUNCH_CHECK();
        return 0;
    }
    dim3 grid(cdiv(p.max_q_len, BQ), p.n_heads, p.n_seg);
    flash_attn_kernel<HD, false><<<grid, NTHREADS, Cfg<HD>::SMEM, st>>>(p);
    PADT_LAUNCH_CHECK();
    return 0;
}

}  // namespace

int launch_attention(const AttnParams& p, cudaStream_t st) {
    if (p.n_seg == 0 || p.max_q_len == 0) return 0;
    PADT_REQUIRE(p.q_row % 8 == 0 && p.k_row % 8 == 0 && p.v_row % 8 == 0 && p.q_head % 8 == 0 && p.k_head % 8 == 0 && p.v_head % 8 == 0,
                 "attention strides must keep 16-byte alignment");
    PADT_REQUIRE(p.o_row % 2 == 0 && p.o_head % 2 == 0, "output strides must be even");
    PADT_REQUIRE(p.n_seg <= 65535 && p.n_heads <= 65535, "grid limits");
    if (attention_tc_eligible(p, p.host_seg)) return launch_attention_tc(p, *p.host_seg, st);
    if (p.head_dim == 80) return launch_hd<80>(p, st);
    if (p.head_dim == 128) return launch_hd<128>(p, st);
    padt_set_error("launch_attention: head_dim must be 80 or 128");
    return -1;
}
