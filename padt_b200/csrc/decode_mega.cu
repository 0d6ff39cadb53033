// Persistent decode-step kernel ("weight-streaming megakernel"): ONE launch per generated token.
//
// Organised for the HBM roofline of the step (6.27 GB per token at 3B, B=1):
//   * grid = one CTA per SM, resident for the whole step; phases (sample | per layer: qkv, attention, o_proj, gate/up, down |
//     logits) are separated by a grid barrier instead of kernel boundaries;
//   * every CTA owns a fixed, contiguous slice of 8-row groups of every projection matrix, so its whole weight read
//     schedule is known up front: a producer warp streams it with 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx)
//     into a shared-memory ring of [8 rows x Kseg] stages, running AHEAD across phase boundaries (while the consumers sit in
//     a grid barrier or in the attention phase the ring keeps filling);
//   * all 16 consumer warps work on the SAME stage, each on its own slice of the K columns: ldmatrix (rows padded by 16 B ->
//     conflict free) + mma.sync m16n8k16 with the weights as the B operand (n = 8 weight rows) and the activation rows as the
//     A operand (m = batch row, <= 8 real rows); a stage is therefore consumed and released in ~100 cycles, so the ring
//     pipeline never drains, and the CUDA cores only run the tiny cross-warp reduction + epilogue per row group.
// Arithmetic: bf16 products accumulated in fp32 by the tensor core, partial sums of the 16 K-slices added in a fixed order,
// then the HF rounding points (bias add / residual add / SiLU*up / logit -> bf16) exactly as in decode.cu.
// Replaces the per-token body of _sample_vision_token / forward_main (padt.py:671-757, 221-229, 268-301).
#include "common.cuh"
#include "kernels.h"
#include "decode_common.cuh"
#include <cstdlib>
#include <vector>

namespace {

constexpr int MK_CW = 16;                     // consumer (tensor-core) warps
constexpr int MK_CTHREADS = MK_CW * 32;
constexpr int MK_STHREADS = MK_CTHREADS + 32; // + 1 epilogue warp: the threads that take part in phase-level barriers
constexpr int MK_THREADS = MK_STHREADS + 32;  // + 1 producer warp
constexpr int MK_NRED = 4;                    // reduction buffers in flight between the consumer warps and the epilogue warp
constexpr int MK_MAXSLOTS = 24;
constexpr int MK_GROUP = 8;                   // weight rows per stage (= n of the mma)
constexpr int MK_MAXGROUPS = 8;               // row groups of one CTA in the small-matrix phases (bias / residual preload)
constexpr int MK_ATTN_SMEM = 2 * DA_TILE * DA_LD * 2 + DA_MAXG * DA_HD * 4 + DA_MAXG * DA_TILE * 4;

struct StageDesc;
struct MegaArgs {
    DecodeParams p;
    const DecodeLayerW* layers;   // device copy
    unsigned int* gbar;           // grid barrier counter, zeroed by a memset node before every launch
    int slot_bytes, n_slots, x_bytes, x_ld;
    const struct StageDesc* stage_tab;   // [sum stages] descriptors, CTA c owns [stage_off[c], stage_off[c+1])
    const int* stage_off;
    int dbg;                      // PADT_MEGA_TIMELINE bits: 1 print cycle breakdown, 2 no compute, 4 no grid barrier, 8 no attention
};

using CSync = NamedSync<1, MK_STHREADS>;


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
__device__ __forceinline__ void ldmatrix_x2(uint32_t& r0, uint32_t& r1, const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];\n" : "=r"(r0), "=r"(r1) : "r"(smem_u32(p)));
}

// ---------------------------------------------------------------- static work partition (identical for producer and consumers)
struct Seg {
    const bf16* img;    // stage image of this CTA's first row group (nullptr: row-major source, one copy per row)
    const bf16* base;   // row `row_lo` of the row-major matrix (used when img == nullptr)
    int row_lo;         // first global output row of this CTA's slice
    int rows;           // valid rows in the slice (the last group may be ragged)
    int groups;         // ceil(rows / 8)
    int K, Kseg, nseg;  // row length, columns per stage, stages per row group
};

// columns per stage: K / nseg (multiple of 16) with the smallest nseg whose 8-row stage fits one ring slot
__host__ __device__ inline int pick_nseg(int K, int slot_bytes) {
    int nseg = 1;
    while (nseg < 64) {
        if (K % (16 * nseg) == 0 && MK_GROUP * ((K / nseg) * 2 + 16) <= slot_bytes) break;
        ++nseg;
    }
    return nseg;
}

__device__ __forceinline__ Seg make_seg(const bf16* W, const bf16* img, int nseg_packed, int row_begin, int n_rows, int K, int cta, int G,
                                         int slot_bytes) {
    Seg s;
    int groups = (n_rows + MK_GROUP - 1) / MK_GROUP;
    int lo = (int)(((long long)groups * cta) / G), hi = (int)(((long long)groups * (cta + 1)) / G);
    int r_lo = lo * MK_GROUP, r_hi = hi * MK_GROUP;
    r_hi = r_hi > n_rows ? n_rows : r_hi;
    s.groups = hi - lo;
    s.rows = r_hi > r_lo ? r_hi - r_lo : 0;
    s.row_lo = row_begin + r_lo;
    s.base = W + (size_t)r_lo * K;
    s.K = K;
    s.nseg = img ? nseg_packed : pick_nseg(K, slot_bytes);
    s.Kseg = K / s.nseg;
    s.img = img ? img + (size_t)lo * s.nseg * MK_GROUP * (s.Kseg + 8) : nullptr;
    return s;
}

enum { SEG_QKV = 0, SEG_O = 1, SEG_GU = 2, SEG_DOWN = 3 };

__device__ __forceinline__ Seg layer_seg(const MegaArgs& a, const DecodeLayerW& w, int which, int cta, int G) {
    const DecodeParams& p = a.p;
    switch (which) {
        case SEG_QKV: return make_seg(w.wqkv, w.iqkv, p.nseg_d, 0, (p.nh + 2 * p.nkv) * p.hd, p.D, cta, G, a.slot_bytes);
        case SEG_O: return make_seg(w.wo, w.io, p.nseg_a, 0, p.D, p.nh * p.hd, cta, G, a.slot_bytes);
        case SEG_GU: return make_seg(w.wgu, w.igu, p.nseg_d, 0, 2 * p.I, p.D, cta, G, a.slot_bytes);
        default: return make_seg(w.wdown, w.idown, p.nseg_i, 0, p.D, p.I, cta, G, a.slot_bytes);
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
// Stage descriptors: the image-backed part of every CTA's weight schedule (layers x {qkv, o, gate/up, down}, then the head
// rows) is identical for every token, so the host flattens it once into a table [cta][stage] of (source, bytes).  The
// producer warp then works 8 stages at a time: lane l owns ring slot l, reads its descriptor (one coalesced 128-byte load
// for the warp), waits for ITS slot to drain and issues ITS bulk copy.  (A single-thread producer loop costs ~0.4 us per
// stage in exposed instruction latency, which is as long as the stage takes to stream.)
struct StageDesc {
    const void* src;
    uint32_t bytes;
    uint32_t pad;
};

__device__ void producer_loop(const MegaArgs& a, uint8_t* ring, uint64_t* full, uint64_t* empty, int cta, int G, int lane) {
    const StageDesc* tab = a.stage_tab + a.stage_off[cta];
    const int n = a.stage_off[cta + 1] - a.stage_off[cta];
    const int ns = a.n_slots;
    if (lane < ns) {
        uint8_t* dst = ring + (size_t)lane * a.slot_bytes;
        uint32_t parity = 0;
        for (int i = lane; i < n; i += ns, parity ^= 1) {
            StageDesc d = tab[i];
            mbar_wait(&empty[lane], parity ^ 1);
            mbar_expect_tx(&full[lane], d.bytes);
            tma_bulk_g2s(dst, d.src, d.bytes, &full[lane]);
        }
    }
    __syncwarp();
    if (lane != 0) return;
    // prototype rows of this request (row-major, one copy per row segment): tiny tail of the stream
    const DecodeParams& p = a.p;
    int plo, phi;
    proto_range(p, plo, phi);
    if (phi <= plo) return;
    Seg s = make_seg(p.protos + (size_t)plo * p.D, nullptr, 0, p.vocab + plo, phi - plo, p.D, cta, G, a.slot_bytes);
    const uint32_t seg_bytes = (uint32_t)s.Kseg * 2, row_stride = seg_bytes + 16;
    uint32_t it = (uint32_t)n;
    for (int g = 0; g < s.groups; ++g) {
        int rows = s.rows - g * MK_GROUP;
        rows = rows > MK_GROUP ? MK_GROUP : rows;
        for (int sg = 0; sg < s.nseg; ++sg, ++it) {
            int slot = it % ns;
            uint32_t par = (it / ns) & 1;
            mbar_wait(&empty[slot], par ^ 1);
            mbar_expect_tx(&full[slot], seg_bytes * rows);
            uint8_t* dst = ring + (size_t)slot * a.slot_bytes;
            const bf16* src = s.base + (size_t)(g * MK_GROUP) * s.K + (size_t)sg * s.Kseg;
            for (int r = 0; r < rows; ++r) tma_bulk_g2s(dst + r * row_stride, src + (size_t)r * s.K, seg_bytes, &full[slot]);
        }
    }
}

// ---------------------------------------------------------------- consumer helpers
// Grid barrier over the consumer threads of all CTAs.  bar.sync orders this CTA's writes before thread 0's gpu-scope
// release-increment (cumulativity); the acquire poll + bar.sync publishes the other CTAs' writes to every thread.
__device__ __forceinline__ void grid_sync(unsigned int* bar, unsigned int& target, int G, int tid, int dbg = 0) {
    CSync()();
    if (dbg & 4) return;
    target += (unsigned int)G;
    if (tid == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
        while (ld_acquire_u32(bar) < target) __nanosleep(32);
    }
    CSync()();
}

// activation rows -> shared memory, row stride x_ld elements (K + 8: the 16-byte pad keeps the A-fragment loads conflict free)
__device__ __forceinline__ void stage_x_plain(const bf16* x, int K, int B, bf16* xs, int x_ld, int tid) {
    if (tid >= MK_CTHREADS) return;
    for (int b = 0; b < B; ++b)
        for (int c = tid * 8; c < K; c += MK_CTHREADS * 8)
            *reinterpret_cast<uint4*>(xs + (size_t)b * x_ld + c) = *reinterpret_cast<const uint4*>(x + (size_t)b * K + c);
}

// RMSNorm prologue, bit-identical to rmsnorm_kernel / the decode.cu prologue (rms_rstd_warp + rms_apply8)
__device__ __forceinline__ void stage_x_rmsnorm(const bf16* h, const bf16* w, float eps, int K, int B, bf16* xs, int x_ld, int tid, int lane,
                                                bf16* copy0, bf16* copy1) {
    if (tid >= MK_CTHREADS) return;
    for (int b = 0; b < B; ++b) {
        const bf16* xr = h + (size_t)b * K;
        float rstd = rms_rstd_warp(xr, K, eps, lane);
        for (int c = tid * 8; c < K; c += MK_CTHREADS * 8) {
            uint4 o = rms_apply8(*reinterpret_cast<const uint4*>(xr + c), *reinterpret_cast<const uint4*>(w + c), rstd);
            *reinterpret_cast<uint4*>(xs + (size_t)b * x_ld + c) = o;
            if (copy0) *reinterpret_cast<uint4*>(copy0 + (size_t)b * K + c) = o;
            if (copy1) *reinterpret_cast<uint4*>(copy1 + (size_t)b * K + c) = o;
        }
    }
}

enum { EP_BIAS = 0, EP_RESID = 1, EP_SWIGLU = 2, EP_LOGITS = 3 };

struct Scratch {
    float* red;                  // [MK_NRED][MK_CW][B][8] partial sums (K slices of the 16 consumer warps)
    uint64_t *red_full, *red_empty;
    float* pre;                  // [8 batch][MK_MAXGROUPS*8] preloaded bias / residual values (epilogue warp only)
};

// Consume one segment from the ring; `it` / `grp` are CTA-wide stage / row-group counters every thread advances identically.
// Consumer warps (0..15): two half-CTAs of 8 warps alternate over the stages (even / odd K segment); inside a stage every warp
// multiplies its own k16-steps (ldmatrix + mma.sync, two independent accumulator chains).  After the last stage of a row group
// a warp drops its 8 partial sums into a reduction buffer and signals an mbarrier -- it never waits for the other warps.
// Epilogue warp (16): waits for the 16 partials of a group, adds them in warp order and applies the HF epilogue.
template <int EP>
__device__ __forceinline__ void consume_seg(const MegaArgs& a, const Seg& s, const uint8_t* ring, uint64_t* full, uint64_t* empty, uint32_t& it,
                                            uint32_t& grp, const bf16* xs, int tid, const bf16* bias, bf16* y, int n_out, const Scratch& sc,
                                            float* best_v, int* best_i, const int* vlo, const int* vhi, long long& t_full) {
    const DecodeParams& p = a.p;
    const int warp = tid >> 5, lane = tid & 31;
    const int B = p.B;
    if (warp == MK_CW) {
        // ================= epilogue warp: lane -> (batch row eb = pass*4 + lane/8, weight row en = lane%8)
        const int en = lane & 7;
        if (EP == EP_BIAS || EP == EP_RESID) {   // preload so that the per-group epilogue never waits on global memory
            for (int pass = 0; pass * 4 < B; ++pass) {
                int eb = pass * 4 + (lane >> 3);
                if (eb < B)
                    for (int gi = 0; gi < s.groups && gi < MK_MAXGROUPS; ++gi) {
                        int r = gi * 8 + en;
                        if (r < s.rows)
                            sc.pre[eb * (MK_MAXGROUPS * 8) + r] = EP == EP_BIAS ? (bias ? bf2f(bias[s.row_lo + r]) : 0.f) : bf2f(y[(size_t)eb * n_out + s.row_lo + r]);
                    }
            }
            __syncwarp();
        }
        for (int gi = 0; gi < s.groups; ++gi, ++grp) {
            const int buf = grp % MK_NRED;
            mbar_wait(&sc.red_full[buf], (grp / MK_NRED) & 1);
            const float* red = sc.red + (size_t)buf * MK_CW * B * 8;
            const int r = gi * 8 + en;
            if (r < s.rows && !(a.dbg & 2)) {
                const int n = s.row_lo + r;
                for (int pass = 0; pass * 4 < B; ++pass) {
                    const int eb = pass * 4 + (lane >> 3);
                    if (eb >= B) break;
                    float sum = 0.f;
#pragma unroll
                    for (int w = 0; w < MK_CW; ++w) sum += red[(w * B + eb) * 8 + en];
                    if (EP == EP_BIAS) {
                        y[(size_t)eb * n_out + n] = f2bf(sum + sc.pre[eb * (MK_MAXGROUPS * 8) + r]);
                    } else if (EP == EP_RESID) {
                        y[(size_t)eb * n_out + n] = f2bf(sc.pre[eb * (MK_MAXGROUPS * 8) + r] + rbf(sum));
                    } else if (EP == EP_SWIGLU) {
                        if ((en & 1) == 0) {   // even row = gate_j, odd row = up_j of the interleaved weight
                            float up = 0.f;
#pragma unroll
                            for (int w = 0; w < MK_CW; ++w) up += red[(w * B + eb) * 8 + en + 1];
                            float gv = rbf(sum), uv = rbf(up);
                            y[(size_t)eb * n_out + (n >> 1)] = f2bf(rbf(silu_f(gv)) * uv);
                        }
                    } else {   // EP_LOGITS: bf16-rounded logit, -inf outside the row's own VRT range, first max wins
                        float l0 = rbf(sum);
                        bool ok = n < p.vocab || (n - p.vocab >= vlo[pass] && n - p.vocab < vhi[pass]);
                        if (ok && (l0 > best_v[pass] || (l0 == best_v[pass] && n < best_i[pass]))) { best_v[pass] = l0; best_i[pass] = n; }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sc.red_empty[buf]);
        }
        it += (uint32_t)s.groups * s.nseg;
        return;
    }
    // ================= consumer warps
    const int g = lane >> 2, t = lane & 3;
    const uint32_t row_stride = (uint32_t)s.Kseg * 2 + 16;
    const int set = warp >> 3, wi = warp & 7;
    const int ksteps = s.Kseg >> 4;
    const int k0 = (ksteps * wi) / 8, k1 = (ksteps * (wi + 1)) / 8;
    for (int gi = 0; gi < s.groups; ++gi, ++grp) {
        float c[4] = {0.f, 0.f, 0.f, 0.f}, c2[4] = {0.f, 0.f, 0.f, 0.f};
        for (int sg = 0; sg < s.nseg; ++sg, ++it) {
            if ((sg & 1) != set && s.nseg > 1) continue;          // the other half owns this stage
            if (s.nseg == 1 && set != 0) continue;
            const int slot = it % a.n_slots;
            const uint32_t parity = (it / a.n_slots) & 1;
            mbar_wait(&full[slot], parity);
            if (!(a.dbg & (2 | 64))) {
                const uint8_t* sb = ring + (size_t)slot * a.slot_bytes;
                const bf16* xrow = xs + (size_t)g * a.x_ld + (size_t)sg * s.Kseg + 2 * t;
                int kk = k0;
#pragma unroll 2
                for (; kk + 1 < k1; kk += 2) {
                    uint32_t b0, b1, b2, b3;
                    ldmatrix_x4(b0, b1, b2, b3, sb + (lane & 7) * row_stride + ((kk + (lane >> 4)) * 16 + ((lane >> 3) & 1) * 8) * 2);
                    uint32_t a0 = 0, a2 = 0, a4 = 0, a6 = 0;
                    if (g < B) {
                        a0 = *reinterpret_cast<const uint32_t*>(xrow + kk * 16);
                        a2 = *reinterpret_cast<const uint32_t*>(xrow + kk * 16 + 8);
                        a4 = *reinterpret_cast<const uint32_t*>(xrow + kk * 16 + 16);
                        a6 = *reinterpret_cast<const uint32_t*>(xrow + kk * 16 + 24);
                    }
                    uint32_t fa[4] = {a0, 0u, a2, 0u}, fb[4] = {a4, 0u, a6, 0u};
                    mma_bf16_16816(c, fa, b0, b1);      // two independent accumulator chains
                    mma_bf16_16816(c2, fb, b2, b3);
                }
                if (kk < k1) {
                    uint32_t b0, b1;
                    ldmatrix_x2(b0, b1, sb + (lane & 7) * row_stride + (kk * 16 + ((lane >> 3) & 1) * 8) * 2);
                    uint32_t a0 = 0, a2 = 0;
                    if (g < B) {
                        a0 = *reinterpret_cast<const uint32_t*>(xrow + kk * 16);
                        a2 = *reinterpret_cast<const uint32_t*>(xrow + kk * 16 + 8);
                    }
                    uint32_t fa[4] = {a0, 0u, a2, 0u};
                    mma_bf16_16816(c, fa, b0, b1);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[slot]);
        }
        float r0 = c[0] + c2[0], r1 = c[1] + c2[1];
        // hand the K-slice partial of this row group to the epilogue warp
        const int buf = grp % MK_NRED;
        mbar_wait(&sc.red_empty[buf], ((grp / MK_NRED) & 1) ^ 1);
        if (g < B) {
            float* red = sc.red + ((size_t)(buf * MK_CW + warp) * B + g) * 8;
            red[2 * t] = r0;
            red[2 * t + 1] = r1;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sc.red_full[buf]);
    }
    (void)t_full;
}

__global__ void __launch_bounds__(MK_THREADS, 1) decode_mega_kernel(MegaArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const DecodeParams& p = a.p;
    uint8_t* ring = smem;
    uint8_t* xa = ring + (size_t)a.n_slots * a.slot_bytes;    // activation rows, aliased with the attention scratch
    bf16* xs = reinterpret_cast<bf16*>(xa);
    uint64_t* full = reinterpret_cast<uint64_t*>(xa + a.x_bytes);
    uint64_t* empty = full + MK_MAXSLOTS;
    uint64_t* red_full = empty + MK_MAXSLOTS;
    uint64_t* red_empty = red_full + MK_NRED;
    float* red = reinterpret_cast<float*>(red_empty + MK_NRED);          // [MK_NRED][MK_CW][B][8]
    __shared__ float s_pre[8 * MK_MAXGROUPS * 8];
    __shared__ float s_bv[64];
    __shared__ int s_bi[64];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = gridDim.x, cta = blockIdx.x;

    if (tid == 0) {
        for (int i = 0; i < a.n_slots; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], MK_CW / 2); }   // half the warps consume a stage
        for (int i = 0; i < MK_NRED; ++i) { mbar_init(&red_full[i], MK_CW); mbar_init(&red_empty[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == MK_CW + 1) {   // producer warp: streams this CTA's weight schedule for the whole step
        producer_loop(a, ring, full, empty, cta, G, lane);
        return;
    }

    Scratch sc;
    sc.red = red; sc.red_full = red_full; sc.red_empty = red_empty;
    sc.pre = s_pre;
    unsigned int target = 0;
    uint32_t it = 0, grp = 0;
    // arg-max state lives in the epilogue warp: lane -> batch rows (lane/8) and 4 + (lane/8)
    float best_v[2] = {-INFINITY, -INFINITY};
    int best_i[2] = {0x7fffffff, 0x7fffffff};
    int vlo[2], vhi[2];
    for (int k = 0; k < 2; ++k) {
        int eb = k * 4 + (lane >> 3);
        vlo[k] = eb < p.B ? p.dev_state[st_vlo(eb)] : 0;
        vhi[k] = eb < p.B ? p.dev_state[st_vhi(eb)] : 0;
    }
    long long t_bar = 0, t_cons = 0, t_stage = 0, t_attn = 0, t_full = 0, t0 = clock64(), tt;
#define TSTART() tt = clock64()
#define TADD(acc) acc += clock64() - tt

    // ---- phase 0: sample the token predicted by the previous forward, log it, gather its embedding row(s) into h
    if (cta == 0) {
        SampleArgs sa;
        sa.amax_val = p.amax_val; sa.amax_idx = p.amax_idx; sa.n_blocks = 0;
        sa.B = p.B; sa.D = p.D; sa.vocab = p.vocab; sa.n_protos = p.n_protos;
        sa.state = p.dev_state; sa.cur_token = p.cur_token; sa.tokens_out = p.tokens_out; sa.argmax_out = p.argmax_out;
        sa.forced = p.forced; sa.max_steps = p.max_steps; sa.eos_ids = p.eos_ids; sa.n_eos = p.n_eos; sa.pad_id = p.pad_id;
        sa.embed = p.embed; sa.protos = p.protos; sa.h = p.h;
        sample_embed_body(sa, tid, MK_STHREADS, CSync());
    }
    TSTART(); grid_sync(a.gbar, target, G, tid, a.dbg); TADD(t_bar);
    const long long t_sample = clock64() - t0;

    const int qkv_n = (p.nh + 2 * p.nkv) * p.hd;
    const size_t layer_kv = (size_t)p.B * p.nkv * p.s_max * p.hd;
    for (int li = 0; li < p.n_layers; ++li) {
        const DecodeLayerW w = a.layers[li];
        // warm L2 with the K/V rows this CTA will attend over after the next barrier (they were written in earlier steps)
        if (tid == 0 && !(a.dbg & 8)) {
            const int items = p.n_split * p.nkv * p.B;
            for (int item = cta; item < items; item += G) {
                int split = item % p.n_split, kvh = (item / p.n_split) % p.nkv, b = item / (p.n_split * p.nkv);
                int cur = p.dev_state[2] + p.dev_state[0] - 1, ks = p.dev_state[st_kvstart(b)];
                int n_keys = cur + 1 - ks;
                int chunk = ((n_keys + p.n_split - 1) / p.n_split + 31) & ~31;
                int kb = ks + split * chunk, ke = kb + chunk;
                ke = ke > cur ? cur : ke;            // the new key is not in the cache yet
                if (ke > kb) {
                    size_t off = (((size_t)li * p.B + b) * p.nkv + kvh) * p.s_max * p.hd + (size_t)kb * p.hd;
                    uint32_t bytes = (uint32_t)(ke - kb) * p.hd * 2;
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.kcache + off), "r"(bytes) : "memory");
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.vcache + off), "r"(bytes) : "memory");
                }
            }
        }
        // (1) RMSNorm -> q/k/v projection + bias
        TSTART();
        stage_x_rmsnorm(p.h, w.ln1, p.eps, p.D, p.B, xs, a.x_ld, tid, lane, nullptr, nullptr);
        CSync()();
        TADD(t_stage); TSTART();
        consume_seg<EP_BIAS>(a, layer_seg(a, w, SEG_QKV, cta, G), ring, full, empty, it, grp, xs, tid, w.bqkv, p.qkv, qkv_n, sc, best_v, best_i, vlo,
                             vhi, t_full);
        TADD(t_cons); TSTART();
        grid_sync(a.gbar, target, G, tid, a.dbg);
        TADD(t_bar); TSTART();
        // (2) M-RoPE + KV append + split-KV attention (+ combine by the last CTA of each (row, kv head))
        {
            DecAttnArgs da;
            da.qkv = p.qkv; da.kcache = p.kcache + li * layer_kv; da.vcache = p.vcache + li * layer_kv;
            da.cos_tab = p.cos_tab; da.sin_tab = p.sin_tab; da.max_pos = p.max_pos; da.sec0 = 0; da.sec1 = 0;
            da.state = p.dev_state; da.part_o = p.part_o; da.part_ml = p.part_ml; da.counters = p.attn_counters;
            da.out = p.act; da.B = p.B; da.nh = p.nh; da.nkv = p.nkv; da.s_max = p.s_max; da.n_split = p.n_split;
            da.scale_log2 = rsqrtf((float)p.hd) * 1.4426950408889634f;
            const int items = p.n_split * p.nkv * p.B;
            for (int item = cta; item < items && !(a.dbg & 8); item += G) {
                int split = item % p.n_split, kvh = (item / p.n_split) % p.nkv, b = item / (p.n_split * p.nkv);
                decode_attn_body(da, split, kvh, b, xa, tid, MK_STHREADS, CSync());
                CSync()();
            }
        }
        TADD(t_attn); TSTART();
        grid_sync(a.gbar, target, G, tid, a.dbg);
        TADD(t_bar); TSTART();
        // (3) o_proj + residual
        stage_x_plain(p.act, p.nh * p.hd, p.B, xs, a.x_ld, tid);
        CSync()();
        TADD(t_stage); TSTART();
        consume_seg<EP_RESID>(a, layer_seg(a, w, SEG_O, cta, G), ring, full, empty, it, grp, xs, tid, nullptr, p.h, p.D, sc, best_v, best_i, vlo, vhi,
                              t_full);
        TADD(t_cons); TSTART();
        grid_sync(a.gbar, target, G, tid, a.dbg);
        TADD(t_bar); TSTART();
        // (4) RMSNorm -> gate/up -> SiLU*up
        stage_x_rmsnorm(p.h, w.ln2, p.eps, p.D, p.B, xs, a.x_ld, tid, lane, nullptr, nullptr);
        CSync()();
        TADD(t_stage); TSTART();
        consume_seg<EP_SWIGLU>(a, layer_seg(a, w, SEG_GU, cta, G), ring, full, empty, it, grp, xs, tid, nullptr, p.act, p.I, sc, best_v, best_i, vlo,
                               vhi, t_full);
        TADD(t_cons); TSTART();
        grid_sync(a.gbar, target, G, tid, a.dbg);
        TADD(t_bar); TSTART();
        // (5) down_proj + residual
        stage_x_plain(p.act, p.I, p.B, xs, a.x_ld, tid);
        CSync()();
        TADD(t_stage); TSTART();
        consume_seg<EP_RESID>(a, layer_seg(a, w, SEG_DOWN, cta, G), ring, full, empty, it, grp, xs, tid, nullptr, p.h, p.D, sc, best_v, best_i, vlo,
                              vhi, t_full);
        TADD(t_cons); TSTART();
        grid_sync(a.gbar, target, G, tid, a.dbg);
        TADD(t_bar);
    }
    const long long t_layers = clock64() - t0;
    // ---- final norm + VRT-extended logits + masked arg-max partial of this CTA
    {
        bf16* log = nullptr;
        if (cta == 0 && p.hidden_steps) {
            int t = p.dev_state[0];
            if (t < p.max_steps) log = p.hidden_steps + (size_t)t * p.B * p.D;
        }
        stage_x_rmsnorm(p.h, p.final_norm, p.eps, p.D, p.B, xs, a.x_ld, tid, lane, cta == 0 ? p.last_hidden : nullptr, log);
        CSync()();
        consume_seg<EP_LOGITS>(a, make_seg(p.head, p.head_img, p.nseg_d, 0, p.vocab, p.D, cta, G, a.slot_bytes), ring, full, empty, it, grp, xs, tid, nullptr, nullptr, 0, sc,
                               best_v, best_i, vlo, vhi, t_full);
        int plo, phi;
        proto_range(p, plo, phi);
        if (phi > plo)
            consume_seg<EP_LOGITS>(a, make_seg(p.protos + (size_t)plo * p.D, nullptr, 0, p.vocab + plo, phi - plo, p.D, cta, G, a.slot_bytes), ring, full,
                                   empty,
                                   it, grp, xs, tid, nullptr, nullptr, 0, sc, best_v, best_i, vlo, vhi, t_full);
        if (warp == MK_CW) {   // epilogue warp publishes its per-lane maxima: entry [batch row][weight-row lane]
            s_bv[lane] = best_v[0]; s_bi[lane] = best_i[0];
            s_bv[32 + lane] = best_v[1]; s_bi[32 + lane] = best_i[1];
        }
        CSync()();
        if (tid < p.B) {
            float bv = -INFINITY; int bi = 0x7fffffff;
            for (int k = 0; k < 8; ++k) {
                float v = s_bv[tid * 8 + k]; int i = s_bi[tid * 8 + k];
                if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
            }
            p.amax_val[(size_t)cta * p.B + tid] = bv;
            p.amax_idx[(size_t)cta * p.B + tid] = bi;
        }
        if (cta == 0 && tid == 0) p.dev_state[3] = G;
    }
    if ((a.dbg & 1) && tid == 0 && (cta == 0 || cta == G / 2 || cta == G - 1))
        printf("[mega cta %d] cycles: total %lld | sample+bar %lld | layers-end %lld | grid-barrier wait %lld | consume %lld (of which waiting for TMA data, warp0: %lld) | x staging %lld | attention %lld\n",
               cta, (long long)(clock64() - t0), t_sample, t_layers, t_bar, t_cons, t_full, t_stage, t_attn);
#undef TSTART
#undef TADD
}

// W [n_rows, K] row-major -> stage image [ceil(n_rows/8)][nseg][8][Kseg + 8] (zero pad, zero rows past n_rows)
__global__ void pack_stage_image_kernel(const bf16* __restrict__ W, int n_rows, int K, int nseg, bf16* __restrict__ out, long long total) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // one 16-byte chunk per thread
    if (i >= total) return;
    const int Kseg = K / nseg, cpr = (Kseg + 8) / 8;                  // chunks per padded row
    int c = (int)(i % cpr);
    long long rowi = i / cpr;
    int r = (int)(rowi % MK_GROUP);
    long long blk = rowi / MK_GROUP;
    int sg = (int)(blk % nseg);
    long long gi = blk / nseg;
    long long row = gi * MK_GROUP + r;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (row < n_rows && c * 8 < Kseg) v = *reinterpret_cast<const uint4*>(W + row * K + (size_t)sg * Kseg + c * 8);
    *reinterpret_cast<uint4*>(out + i * 8) = v;
}

}  // namespace

static int stage_cap() { return 24 * 1024; }
int decode_mega_nseg(int K) { return pick_nseg(K, stage_cap()); }

size_t decode_mega_image_elems(int n_rows, int K) {
    int nseg = decode_mega_nseg(K);
    size_t groups = (n_rows + MK_GROUP - 1) / MK_GROUP;
    return groups * nseg * MK_GROUP * (size_t)(K / nseg + 8);
}

int launch_pack_stage_image(const bf16* W, int n_rows, int K, bf16* out, cudaStream_t st) {
    int nseg = decode_mega_nseg(K);
    PADT_REQUIRE(nseg < 64 && K % (16 * nseg) == 0, "row length cannot be split into stage segments");
    long long total = (long long)decode_mega_image_elems(n_rows, K) / 8;
    pack_stage_image_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(W, n_rows, K, nseg, out, total);
    PADT_LAUNCH_CHECK();
    return 0;
}

// Plans shared memory for the persistent kernel; returns 0 and fills the outputs, or -1 when the model does not fit.
int decode_mega_plan(const DecodeParams& p, int* slot_bytes, int* n_slots, int* x_bytes, size_t* smem_bytes) {
    if (p.B > 8) return -1;
    int kmax = p.I > p.D ? p.I : p.D;
    kmax = kmax > p.nh * p.hd ? kmax : p.nh * p.hd;
    int xb = p.B * (kmax + 8) * 2;
    xb = xb > MK_ATTN_SMEM ? xb : MK_ATTN_SMEM;
    xb = (xb + 127) / 128 * 128;
    // one slot = 8 rows x Kseg columns (+16 B pad per row); stages of at most 24 KB
    int sb = 0;
    const int ks[3] = {p.D, p.nh * p.hd, p.I};
    for (int i = 0; i < 3; ++i) {
        int nseg = pick_nseg(ks[i], stage_cap());
        if (nseg >= 64) return -1;
        int bytes = MK_GROUP * ((ks[i] / nseg) * 2 + 16);
        sb = bytes > sb ? bytes : sb;
    }
    sb = (sb + 127) / 128 * 128;
    const int budget = 227 * 1024 - 6 * 1024;    // static shared memory (reduction scratch 8 KB, preload 2 KB, arg-max) + slack
    int fixed = xb + (2 * MK_MAXSLOTS + 2 * MK_NRED) * 8 + MK_NRED * MK_CW * p.B * 8 * 4;
    int ns = (budget - fixed) / sb;
    ns = ns > MK_MAXSLOTS ? MK_MAXSLOTS : ns;
    if (ns < 3) return -1;
    *slot_bytes = sb; *n_slots = ns; *x_bytes = xb;
    *smem_bytes = (size_t)ns * sb + fixed;
    return 0;
}

// Host mirror of make_seg/layer_seg: flattens the image-backed stages of every CTA (same order as the consumers walk them).
int decode_mega_build_table(const DecodeParams& p, const DecodeLayerW* layers_host, int G, std::vector<uint64_t>& tab, std::vector<int>& off) {
    tab.clear();
    off.assign(G + 1, 0);
    auto add_seg = [&](const bf16* img, int n_rows, int K) {
        // appended per CTA below
        (void)img; (void)n_rows; (void)K;
    };
    (void)add_seg;
    for (int c = 0; c < G; ++c) {
        off[c] = (int)(tab.size() / 2);
        auto emit = [&](const bf16* img, int n_rows, int K) {
            int nseg = decode_mega_nseg(K);
            int Kseg = K / nseg;
            int groups = (n_rows + MK_GROUP - 1) / MK_GROUP;
            long long lo = ((long long)groups * c) / G, hi = ((long long)groups * (c + 1)) / G;
            uint32_t bytes = MK_GROUP * ((uint32_t)Kseg * 2 + 16);
            for (long long g = lo; g < hi; ++g)
                for (int sg = 0; sg < nseg; ++sg) {
                    const uint8_t* src = reinterpret_cast<const uint8_t*>(img) + (size_t)(g * nseg + sg) * bytes;
                    tab.push_back((uint64_t)(uintptr_t)src);
                    tab.push_back((uint64_t)bytes);
                }
        };
        for (int li = 0; li < p.n_layers; ++li) {
            const DecodeLayerW& w = layers_host[li];
            emit(w.iqkv, (p.nh + 2 * p.nkv) * p.hd, p.D);
            emit(w.io, p.D, p.nh * p.hd);
            emit(w.igu, 2 * p.I, p.D);
            emit(w.idown, p.D, p.I);
        }
        emit(p.head_img, p.vocab, p.D);
    }
    off[G] = (int)(tab.size() / 2);
    return 0;
}

int launch_decode_mega(const DecodeParams& p, const DecodeLayerW* layers_dev, const void* stage_tab, const int* stage_off, unsigned int* gbar,
                       int n_sms, cudaStream_t st) {
    PADT_REQUIRE(p.hd == DA_HD && p.nh % p.nkv == 0 && p.nh / p.nkv <= DA_MAXG, "head layout");
    PADT_REQUIRE(p.D % 16 == 0 && p.I % 16 == 0, "row lengths must be multiples of 16");
    PADT_REQUIRE(p.n_split <= 64, "at most 64 KV splits");
    MegaArgs a;
    a.p = p; a.layers = layers_dev; a.gbar = gbar;
    a.stage_tab = reinterpret_cast<const StageDesc*>(stage_tab); a.stage_off = stage_off;
    a.dbg = getenv("PADT_MEGA_TIMELINE") ? atoi(getenv("PADT_MEGA_TIMELINE")) : 0;
    size_t smem;
    if (decode_mega_plan(p, &a.slot_bytes, &a.n_slots, &a.x_bytes, &smem)) {
        padt_set_error("persistent decode kernel: configuration does not fit in shared memory");
        return -1;
    }
    int kmax = p.I > p.D ? p.I : p.D;
    kmax = kmax > p.nh * p.hd ? kmax : p.nh * p.hd;
    a.x_ld = kmax + 8;
    // bias / residual preload covers MK_MAXGROUPS row groups per CTA in the q|k|v, o_proj and down_proj phases
    const int qkv_groups = ((p.nh + 2 * p.nkv) * p.hd / MK_GROUP + n_sms - 1) / n_sms + 1;
    PADT_REQUIRE(qkv_groups <= MK_MAXGROUPS && (p.D / MK_GROUP + n_sms - 1) / n_sms + 1 <= MK_MAXGROUPS,
                 "too many row groups per CTA for the preload buffer");
    PADT_REQUIRE(((uintptr_t)p.head & 15) == 0 && (p.protos == nullptr || ((uintptr_t)p.protos & 15) == 0), "TMA bulk copies need 16-byte aligned rows");
    static size_t configured = 0;
    if (smem > configured) {
        PADT_CHECK_CUDA(cudaFuncSetAttribute(decode_mega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    PADT_CHECK_CUDA(cudaMemsetAsync(a.gbar, 0, sizeof(unsigned int), st));
    void* args[] = {(void*)&a};
    PADT_CHECK_CUDA(cudaLaunchCooperativeKernel((const void*)decode_mega_kernel, dim3(n_sms), dim3(MK_THREADS), args, smem, st));
    ++g_padt_launches;
    return 0;
}
