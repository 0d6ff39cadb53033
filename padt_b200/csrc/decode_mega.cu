// Persistent decode-step kernel ("weight-streaming megakernel"): ONE launch per generated token.
//
// Same arithmetic as decode.cu (shared device code in decode_common.cuh, identical chunk/FMA/reduction order, so the two
// paths agree bit for bit) but organised for the HBM roofline of the step (6.27 GB per token at 3B):
//   * grid = one CTA per SM, resident for the whole step; phases (sample | per layer: qkv, attention, o_proj, gate/up, down |
//     logits) are separated by a grid barrier instead of kernel boundaries;
//   * every CTA owns a fixed, contiguous slice of the rows of every projection matrix, so its whole weight read schedule is
//     known up front: a producer warp streams it with 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx) into a
//     shared-memory ring, running AHEAD across phase boundaries (while the consumers sit in a grid barrier or in the
//     attention phase the ring keeps filling, so HBM never idles waiting for a dependent activation);
//   * 8 consumer warps compute row dot-products straight from the ring against the activation vector held in shared memory.
// Replaces the per-token body of _sample_vision_token / forward_main (padt.py:671-757, 221-229, 268-301).
#include "common.cuh"
#include "kernels.h"
#include "decode_common.cuh"

namespace {

constexpr int MK_CW = 8;                      // consumer warps
constexpr int MK_CTHREADS = MK_CW * 32;
constexpr int MK_THREADS = MK_CTHREADS + 32;  // + 1 producer warp
constexpr int MK_MAXSLOTS = 12;
constexpr int MK_ATTN_SMEM = 2 * DA_TILE * DA_LD * 2 + DA_MAXG * DA_HD * 4 + DA_MAXG * DA_TILE * 4;

struct MegaArgs {
    DecodeParams p;
    const DecodeLayerW* layers;   // device copy
    unsigned int* gbar;           // grid barrier counter, zeroed by a memset node before every launch
    int slot_bytes, n_slots, x_bytes;
};

using CSync = NamedSync<1, MK_CTHREADS>;

// ---------------------------------------------------------------- mbarrier / TMA-bulk primitives
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(b)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// ---------------------------------------------------------------- static work partition (identical for producer and consumers)
struct Seg {
    const bf16* base;   // first row of this CTA's slice
    int row_lo;         // global index of that row
    int rows;           // rows in the slice (multiple of `unit`)
    int K, R;           // row length, rows per ring stage
    __device__ __forceinline__ int n_stages() const { return (rows + R - 1) / R; }
};

__device__ __forceinline__ Seg make_seg(const bf16* W, int row_begin, int n_rows, int K, int unit, int cta, int G, int slot_bytes) {
    Seg s;
    int units = n_rows / unit;
    int lo = (int)(((long long)units * cta) / G), hi = (int)(((long long)units * (cta + 1)) / G);
    s.row_lo = row_begin + lo * unit;
    s.rows = (hi - lo) * unit;
    s.base = W + (size_t)(lo * unit) * K;
    s.K = K;
    int R = slot_bytes / (K * 2);
    R = R < 1 ? 1 : R;
    if (unit == 2) R &= ~1;
    s.R = R;
    return s;
}

enum { SEG_QKV = 0, SEG_O = 1, SEG_GU = 2, SEG_DOWN = 3 };

__device__ __forceinline__ Seg layer_seg(const MegaArgs& a, const DecodeLayerW& w, int which, int cta, int G) {
    const DecodeParams& p = a.p;
    switch (which) {
        case SEG_QKV: return make_seg(w.wqkv, 0, (p.nh + 2 * p.nkv) * p.hd, p.D, 1, cta, G, a.slot_bytes);
        case SEG_O: return make_seg(w.wo, 0, p.D, p.nh * p.hd, 1, cta, G, a.slot_bytes);
        case SEG_GU: return make_seg(w.wgu, 0, 2 * p.I, p.D, 2, cta, G, a.slot_bytes);
        default: return make_seg(w.wdown, 0, p.D, p.I, 1, cta, G, a.slot_bytes);
    }
}

// prototype rows any batch row may emit: [lo, hi) over the union of the per-row ranges (contiguous for rank-local batches)
__device__ __forceinline__ void proto_range(const DecodeParams& p, int& lo, int& hi) {
    lo = 0x7fffffff; hi = 0;
    for (int b = 0; b < p.B; ++b) {
        int l = p.dev_state[st_vlo(b)], h = p.dev_state[st_vhi(b)];
        lo = l < lo ? l : lo;
        hi = h > hi ? h : hi;
    }
    if (hi < lo) { lo = 0; hi = 0; }
}

// ---------------------------------------------------------------- producer: stream this CTA's weight slices through the ring
__device__ void producer_loop(const MegaArgs& a, uint8_t* ring, uint64_t* full, uint64_t* empty, int cta, int G) {
    uint32_t it = 0;
    auto stream = [&](const Seg& s) {
        const int ns = s.n_stages();
        for (int st = 0; st < ns; ++st, ++it) {
            int slot = it % a.n_slots;
            uint32_t parity = (it / a.n_slots) & 1;
            mbar_wait(&empty[slot], parity ^ 1);
            int rows = s.rows - st * s.R;
            rows = rows > s.R ? s.R : rows;
            uint32_t bytes = (uint32_t)rows * s.K * 2;
            mbar_expect_tx(&full[slot], bytes);
            tma_bulk_g2s(ring + (size_t)slot * a.slot_bytes, s.base + (size_t)st * s.R * s.K, bytes, &full[slot]);
        }
    };
    const DecodeParams& p = a.p;
    for (int li = 0; li < p.n_layers; ++li) {
        const DecodeLayerW w = a.layers[li];
        for (int k = 0; k < 4; ++k) stream(layer_seg(a, w, k, cta, G));
    }
    stream(make_seg(p.head, 0, p.vocab, p.D, 1, cta, G, a.slot_bytes));
    int plo, phi;
    proto_range(p, plo, phi);
    if (phi > plo) stream(make_seg(p.protos + (size_t)plo * p.D, p.vocab + plo, phi - plo, p.D, 1, cta, G, a.slot_bytes));
}

// ---------------------------------------------------------------- consumer helpers
__device__ __forceinline__ void grid_sync(unsigned int* bar, unsigned int& target, int G, int tid) {
    CSync()();
    target += (unsigned int)G;
    if (tid == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        while (ld_acquire_u32(bar) < target) { }
        __threadfence();
    }
    CSync()();
}

// one weight row (in shared memory) against NB activation rows; same chunk / FMA order as warp_rows2 in decode_common.cuh
template <int NB>
__device__ __forceinline__ void warp_row_smem(const bf16* r0, int K, const bf16* xs, int lane, float* acc) {
#pragma unroll
    for (int b = 0; b < NB; ++b) acc[b] = 0.f;
    const int nch = K >> 3;
    for (int c = lane; c < nch; c += 32) {
        uint4 w = *reinterpret_cast<const uint4*>(r0 + (size_t)c * 8);
        dot8<NB>(w, xs, K, c, acc);
    }
#pragma unroll
    for (int b = 0; b < NB; ++b) acc[b] = warp_sum(acc[b]);
}

template <int NB>
__device__ __forceinline__ void stage_x_plain(const bf16* x, int K, int B, bf16* xs, int tid) {
    for (int b = 0; b < NB; ++b)
        for (int c = tid * 8; c < K; c += MK_CTHREADS * 8)
            *reinterpret_cast<uint4*>(xs + (size_t)b * K + c) = b < B ? *reinterpret_cast<const uint4*>(x + (size_t)b * K + c) : make_uint4(0, 0, 0, 0);
}

// RMSNorm prologue, bit-identical to rmsnorm_kernel / the decode.cu prologue (rms_rstd_warp + rms_apply8)
template <int NB>
__device__ __forceinline__ void stage_x_rmsnorm(const bf16* h, const bf16* w, float eps, int K, int B, bf16* xs, int tid, int lane, bf16* copy0,
                                                bf16* copy1) {
    for (int b = 0; b < NB; ++b) {
        if (b < B) {
            const bf16* xr = h + (size_t)b * K;
            float rstd = rms_rstd_warp(xr, K, eps, lane);
            for (int c = tid * 8; c < K; c += MK_CTHREADS * 8) {
                uint4 o = rms_apply8(*reinterpret_cast<const uint4*>(xr + c), *reinterpret_cast<const uint4*>(w + c), rstd);
                *reinterpret_cast<uint4*>(xs + (size_t)b * K + c) = o;
                if (copy0) *reinterpret_cast<uint4*>(copy0 + (size_t)b * K + c) = o;
                if (copy1) *reinterpret_cast<uint4*>(copy1 + (size_t)b * K + c) = o;
            }
        } else {
            for (int c = tid * 8; c < K; c += MK_CTHREADS * 8) *reinterpret_cast<uint4*>(xs + (size_t)b * K + c) = make_uint4(0, 0, 0, 0);
        }
    }
}

enum { EP_BIAS = 0, EP_RESID = 1, EP_SWIGLU = 2, EP_LOGITS = 3 };

struct Best {
    float v[8];
    int i[8];
};

// consume one segment from the ring; `it` is the CTA-wide stage counter shared (by construction) with the producer
template <int NB, int EP>
__device__ __forceinline__ void consume_seg(const MegaArgs& a, const Seg& s, const uint8_t* ring, uint64_t* full, uint64_t* empty, uint32_t& it,
                                            const bf16* xs, int warp, int lane, const bf16* bias, bf16* y, int n_out, Best& best,
                                            const int* vlo, const int* vhi) {
    const DecodeParams& p = a.p;
    constexpr int unit = (EP == EP_SWIGLU) ? 2 : 1;
    const int U = s.R / unit;
    const int ns = s.n_stages();
    for (int st = 0; st < ns; ++st, ++it) {
        const int slot = it % a.n_slots;
        const uint32_t parity = (it / a.n_slots) & 1;
        mbar_wait(&full[slot], parity);
        int rows = s.rows - st * s.R;
        rows = rows > s.R ? s.R : rows;
        const int units = rows / unit;
        const bf16* sbase = reinterpret_cast<const bf16*>(ring + (size_t)slot * a.slot_bytes);
        for (int ui = 0; ui < units; ++ui) {
            if (((st * U + ui) & (MK_CW - 1)) != warp) continue;
            const int n = s.row_lo + st * s.R + ui * unit;   // global output row
            const bf16* r0 = sbase + (size_t)ui * unit * s.K;
            float acc0[NB], acc1[NB];
            warp_row_smem<NB>(r0, s.K, xs, lane, acc0);
            if (EP == EP_SWIGLU) warp_row_smem<NB>(r0 + s.K, s.K, xs, lane, acc1);
            if (EP == EP_LOGITS) {
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    float l0 = rbf(acc0[b]);
                    bool ok = n < p.vocab || (n - p.vocab >= vlo[b] && n - p.vocab < vhi[b]);
                    if (ok && (l0 > best.v[b] || (l0 == best.v[b] && n < best.i[b]))) { best.v[b] = l0; best.i[b] = n; }
                }
            } else if (lane == 0) {
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    if (b >= p.B) break;
                    if (EP == EP_BIAS) {
                        y[(size_t)b * n_out + n] = f2bf(acc0[b] + (bias ? bf2f(bias[n]) : 0.f));
                    } else if (EP == EP_RESID) {
                        bf16* h0 = y + (size_t)b * n_out + n;
                        h0[0] = f2bf(bf2f(h0[0]) + rbf(acc0[b]));
                    } else if (EP == EP_SWIGLU) {
                        float g = rbf(acc0[b]), u = rbf(acc1[b]);
                        y[(size_t)b * n_out + (n >> 1)] = f2bf(rbf(silu_f(g)) * u);
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[slot]);
    }
}

template <int NB>
__global__ void __launch_bounds__(MK_THREADS, 1) decode_mega_kernel(MegaArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const DecodeParams& p = a.p;
    uint8_t* ring = smem;
    uint8_t* xa = ring + (size_t)a.n_slots * a.slot_bytes;    // activation rows, aliased with the attention scratch
    bf16* xs = reinterpret_cast<bf16*>(xa);
    uint64_t* full = reinterpret_cast<uint64_t*>(xa + a.x_bytes);
    uint64_t* empty = full + MK_MAXSLOTS;
    __shared__ float s_val[MK_CW][NB];
    __shared__ int s_idx[MK_CW][NB];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = gridDim.x, cta = blockIdx.x;

    if (tid == 0) {
        for (int i = 0; i < a.n_slots; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], MK_CW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == MK_CW) {   // producer warp: one lane streams weights for the whole step
        if (lane == 0) producer_loop(a, ring, full, empty, cta, G);
        return;
    }

    unsigned int target = 0;
    uint32_t it = 0;
    Best best;
    int vlo[NB], vhi[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) { best.v[b] = -INFINITY; best.i[b] = 0x7fffffff; vlo[b] = b < p.B ? p.dev_state[st_vlo(b)] : 0; vhi[b] = b < p.B ? p.dev_state[st_vhi(b)] : 0; }

    // ---- phase 0: sample the token predicted by the previous forward, log it, gather its embedding row(s) into h
    if (cta == 0) {
        SampleArgs sa;
        sa.amax_val = p.amax_val; sa.amax_idx = p.amax_idx; sa.n_blocks = 0;
        sa.B = p.B; sa.D = p.D; sa.vocab = p.vocab; sa.n_protos = p.n_protos;
        sa.state = p.dev_state; sa.cur_token = p.cur_token; sa.tokens_out = p.tokens_out; sa.argmax_out = p.argmax_out;
        sa.forced = p.forced; sa.max_steps = p.max_steps; sa.eos_ids = p.eos_ids; sa.n_eos = p.n_eos; sa.pad_id = p.pad_id;
        sa.embed = p.embed; sa.protos = p.protos; sa.h = p.h;
        sample_embed_body(sa, tid, MK_CTHREADS, CSync());
    }
    grid_sync(a.gbar, target, G, tid);

    const int qkv_n = (p.nh + 2 * p.nkv) * p.hd;
    const size_t layer_kv = (size_t)p.B * p.nkv * p.s_max * p.hd;
    for (int li = 0; li < p.n_layers; ++li) {
        const DecodeLayerW w = a.layers[li];
        // (1) RMSNorm -> q/k/v projection + bias
        stage_x_rmsnorm<NB>(p.h, w.ln1, p.eps, p.D, p.B, xs, tid, lane, nullptr, nullptr);
        CSync()();
        consume_seg<NB, EP_BIAS>(a, layer_seg(a, w, SEG_QKV, cta, G), ring, full, empty, it, xs, warp, lane, w.bqkv, p.qkv, qkv_n, best, vlo, vhi);
        grid_sync(a.gbar, target, G, tid);
        // (2) M-RoPE + KV append + split-KV attention (+ combine by the last CTA of each (row, kv head))
        {
            DecAttnArgs da;
            da.qkv = p.qkv; da.kcache = p.kcache + li * layer_kv; da.vcache = p.vcache + li * layer_kv;
            da.cos_tab = p.cos_tab; da.sin_tab = p.sin_tab; da.max_pos = p.max_pos; da.sec0 = 0; da.sec1 = 0;
            da.state = p.dev_state; da.part_o = p.part_o; da.part_ml = p.part_ml; da.counters = p.attn_counters;
            da.out = p.act; da.B = p.B; da.nh = p.nh; da.nkv = p.nkv; da.s_max = p.s_max; da.n_split = p.n_split;
            da.scale_log2 = rsqrtf((float)p.hd) * 1.4426950408889634f;
            const int items = p.n_split * p.nkv * p.B;
            for (int item = cta; item < items; item += G) {
                int split = item % p.n_split, kvh = (item / p.n_split) % p.nkv, b = item / (p.n_split * p.nkv);
                decode_attn_body(da, split, kvh, b, xa, tid, CSync());
                CSync()();
            }
        }
        grid_sync(a.gbar, target, G, tid);
        // (3) o_proj + residual
        stage_x_plain<NB>(p.act, p.nh * p.hd, p.B, xs, tid);
        CSync()();
        consume_seg<NB, EP_RESID>(a, layer_seg(a, w, SEG_O, cta, G), ring, full, empty, it, xs, warp, lane, nullptr, p.h, p.D, best, vlo, vhi);
        grid_sync(a.gbar, target, G, tid);
        // (4) RMSNorm -> gate/up -> SiLU*up
        stage_x_rmsnorm<NB>(p.h, w.ln2, p.eps, p.D, p.B, xs, tid, lane, nullptr, nullptr);
        CSync()();
        consume_seg<NB, EP_SWIGLU>(a, layer_seg(a, w, SEG_GU, cta, G), ring, full, empty, it, xs, warp, lane, nullptr, p.act, p.I, best, vlo, vhi);
        grid_sync(a.gbar, target, G, tid);
        // (5) down_proj + residual
        stage_x_plain<NB>(p.act, p.I, p.B, xs, tid);
        CSync()();
        consume_seg<NB, EP_RESID>(a, layer_seg(a, w, SEG_DOWN, cta, G), ring, full, empty, it, xs, warp, lane, nullptr, p.h, p.D, best, vlo, vhi);
        grid_sync(a.gbar, target, G, tid);
    }
    // ---- final norm + VRT-extended logits + masked arg-max partial of this CTA
    {
        bf16* log = nullptr;
        if (cta == 0 && p.hidden_steps) {
            int t = p.dev_state[0];
            if (t < p.max_steps) log = p.hidden_steps + (size_t)t * p.B * p.D;
        }
        stage_x_rmsnorm<NB>(p.h, p.final_norm, p.eps, p.D, p.B, xs, tid, lane, cta == 0 ? p.last_hidden : nullptr, log);
        CSync()();
        consume_seg<NB, EP_LOGITS>(a, make_seg(p.head, 0, p.vocab, p.D, 1, cta, G, a.slot_bytes), ring, full, empty, it, xs, warp, lane, nullptr,
                                   nullptr, 0, best, vlo, vhi);
        int plo, phi;
        proto_range(p, plo, phi);
        if (phi > plo)
            consume_seg<NB, EP_LOGITS>(a, make_seg(p.protos + (size_t)plo * p.D, p.vocab + plo, phi - plo, p.D, 1, cta, G, a.slot_bytes), ring, full,
                                       empty, it, xs, warp, lane, nullptr, nullptr, 0, best, vlo, vhi);
        if (lane == 0) {
#pragma unroll
            for (int b = 0; b < NB; ++b) { s_val[warp][b] = best.v[b]; s_idx[warp][b] = best.i[b]; }
        }
        CSync()();
        if (tid < p.B) {
            float bv = -INFINITY; int bi = 0x7fffffff;
            for (int k = 0; k < MK_CW; ++k) {
                float v = s_val[k][tid]; int i = s_idx[k][tid];
                if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
            }
            p.amax_val[(size_t)cta * p.B + tid] = bv;
            p.amax_idx[(size_t)cta * p.B + tid] = bi;
        }
        if (cta == 0 && tid == 0) p.dev_state[3] = G;
    }
}

template <int NB>
int launch_nb(const MegaArgs& a, int grid, size_t smem, cudaStream_t st) {
    static size_t configured = 0;
    if (smem > configured) {
        PADT_CHECK_CUDA(cudaFuncSetAttribute(decode_mega_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    PADT_CHECK_CUDA(cudaMemsetAsync(a.gbar, 0, sizeof(unsigned int), st));
    void* args[] = {(void*)&a};
    PADT_CHECK_CUDA(cudaLaunchCooperativeKernel((const void*)decode_mega_kernel<NB>, dim3(grid), dim3(MK_THREADS), args, smem, st));
    ++g_padt_launches;
    return 0;
}

}  // namespace

// Plans shared memory for the persistent kernel; returns 0 and fills the outputs, or -1 when the model does not fit.
int decode_mega_plan(const DecodeParams& p, int* slot_bytes, int* n_slots, int* x_bytes, size_t* smem_bytes) {
    const int NB = p.B == 1 ? 1 : (p.B == 2 ? 2 : 4);
    int kmax = p.I > p.D ? p.I : p.D;
    kmax = kmax > p.nh * p.hd ? kmax : p.nh * p.hd;
    int xb = NB * kmax * 2;
    xb = xb > MK_ATTN_SMEM ? xb : MK_ATTN_SMEM;
    xb = (xb + 127) / 128 * 128;
    int sb = p.I * 2;
    sb = sb < 16384 ? 16384 : sb;
    sb = (sb + 127) / 128 * 128;
    const int budget = 227 * 1024 - 1024;   // static smem (reduction scratch) + slack
    int fixed = xb + 2 * MK_MAXSLOTS * 8;
    int ns = (budget - fixed) / sb;
    ns = ns > MK_MAXSLOTS ? MK_MAXSLOTS : ns;
    if (ns < 2 || p.B > 4) return -1;
    *slot_bytes = sb; *n_slots = ns; *x_bytes = xb;
    *smem_bytes = (size_t)ns * sb + fixed;
    return 0;
}

int launch_decode_mega(const DecodeParams& p, const DecodeLayerW* layers_dev, unsigned int* gbar, int n_sms, cudaStream_t st) {
    PADT_REQUIRE(p.hd == DA_HD && p.nh % p.nkv == 0 && p.nh / p.nkv <= DA_MAXG, "head layout");
    PADT_REQUIRE(p.D % 8 == 0 && p.I % 8 == 0, "row lengths must be multiples of 8");
    MegaArgs a;
    a.p = p; a.layers = layers_dev; a.gbar = gbar;
    size_t smem;
    if (decode_mega_plan(p, &a.slot_bytes, &a.n_slots, &a.x_bytes, &smem)) {
        padt_set_error("persistent decode kernel: configuration does not fit in shared memory");
        return -1;
    }
    PADT_REQUIRE(((uintptr_t)p.head & 15) == 0 && (p.protos == nullptr || ((uintptr_t)p.protos & 15) == 0), "TMA bulk copies need 16-byte aligned rows");
    if (p.B == 1) return launch_nb<1>(a, n_sms, smem, st);
    if (p.B == 2) return launch_nb<2>(a, n_sms, smem, st);
    return launch_nb<4>(a, n_sms, smem, st);
}
