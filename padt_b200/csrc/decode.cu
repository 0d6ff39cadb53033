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
#include "decode_common.cuh"

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
    int* state;           // device state block
    float* amax_val; int* amax_idx;
    bf16* last_hidden;    // PRO_RMSNORM + GE_LOGITS: normalised rows also stored here
    bf16* hidden_steps;   // optional log [T,B,K]
    int max_steps;
    bf16* logits_out; int logits_ld;   // optional [B, logits_ld] masked bf16 logits (sampling / logits processors)
};

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
                if (a.logits_out && lane == 0 && b < a.B) {
                    a.logits_out[(size_t)b * a.logits_ld + n0] = f2bf(ok0 ? l0 : -INFINITY);
                    if (has1) a.logits_out[(size_t)b * a.logits_ld + n1] = f2bf(ok1 ? l1 : -INFINITY);
                }
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
        if (blockIdx.x == 0 && tid == 0) a.state[3] = gridDim.x;   // number of partials the sampler must reduce
    }
}

__global__ void __launch_bounds__(256) sample_embed_kernel(SampleArgs a) {
    sample_embed_body(a, threadIdx.x, blockDim.x, BlockSync());
}

__global__ void __launch_bounds__(DA_THREADS) decode_attn_kernel(DecAttnArgs a) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    decode_attn_body<true>(a, blockIdx.x, blockIdx.y, blockIdx.z, smem_raw, threadIdx.x, DA_THREADS, BlockSync());
}

template <int NB, int PRO, int EPI>
int launch_gemv_t(const GemvArgs& a, int grid, cudaStream_t st) {
    size_t smem = (size_t)NB * a.K * sizeof(bf16);
    static size_t configured_dev[64] = {0};
    size_t& configured = configured_dev[padt_dev_slot()];
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
    a.logits_out = p.logits_out; a.logits_ld = p.logits_ld;
    PADT_REQUIRE(p.n_amax_blocks >= gemv_grid(a.N), "arg-max partial buffer too small");
    return launch_gemv<PRO_RMSNORM, GE_LOGITS>(a, gemv_grid(a.N), st);
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
