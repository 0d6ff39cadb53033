// Persistent decode-step kernel ("weight-streaming megakernel"): ONE launch per generated token.
//
// Organised for the HBM roofline of the step (6.27 GB per token at 3B, B=1); DESIGN.md §4.1 has the measurements.
//   * grid = one CTA per SM, resident for the whole step: 16 tensor-core consumer warps, 1 epilogue warp, 1 TMA producer warp,
//     1 L2 prefetch warp.  Phases (sample | per layer: qkv, attention, o_proj, gate/up, down | head rows | prototype rows) run in
//     ONE phase loop.  There is NO grid barrier between them: every vector one phase hands to the next (residual stream, q|k|v,
//     split-KV partials, attention output, SwiGLU activations) travels as 8-byte "LL" words {phase epoch, payload} written with
//     relaxed stores and POLLED by the consumers' staging loops (flag and data share one single-copy-atomic word: no fence, no
//     counter, no second round trip).  The 5 barriers per layer of the previous version (1.6-2.1 us each + arrival skew + the
//     post-barrier staging read) were the largest term of the step; the sample phase runs redundantly in every CTA;
//   * every CTA owns a fixed slice of every projection matrix, so its whole weight read schedule is known up front (a host-built
//     descriptor table): producer lane l streams the stages of ring slot l with 1-D TMA bulk copies (cp.async.bulk + mbarrier
//     complete_tx, L2 evict_first) into a shared-memory ring of [8 rows x Kseg] stages, gated only by free slots, i.e. running
//     AHEAD across phase boundaries; the prefetch warp walks the same table further ahead and pulls stages into L2, so the stream
//     continues while the consumers sit in a grid barrier or in the attention phase and the ring is full;
//   * two half-CTAs of 8 warps alternate over the stages (the ring has an EVEN slot count, so a slot always belongs to one half);
//     inside a stage every warp owns a K slice: ldmatrix.x4 on the padded stage (conflict free) + mma.sync m16n8k16 with the
//     weights as the B operand (n = 8 weight rows) and the batch rows as the A operand (<= 8 real rows).  Shared-memory addresses
//     are computed once per phase and the activation fragments of K = 2048 phases stay in registers, so a stage costs
//     4 x (LDSM.x4 + 2 HMMA) per warp; the epilogue warp adds the 16 K-slice partials of a row group in fixed order and applies
//     the HF rounding points (bias add / residual add / SiLU*up / bf16 logit + masked arg-max) exactly as decode.cu does;
//   * attention: (row, kv head, split) items; the item's K/V tile is copied to shared memory BEFORE the barrier that publishes
//     q/k/v; the CTAs of a kv head meet on a monotonic counter and each combines a slice of the outputs over all splits;
//   * down_proj is distributed by stages, not row groups (balanced); a row group shared by two CTAs is completed by the CTA that
//     owns its first K segments from the neighbour's published fp32 partial ({epoch, value} words, fixed order).
// Replaces the per-token body of _sample_vision_token / forward_main (padt.py:671-757, 221-229, 268-301).
#include "common.cuh"
#include "kernels.h"
#include "decode_common.cuh"
#include <cstdlib>
#include <vector>

namespace {

constexpr int MK_CW = 8;                      // consumer (tensor-core) warps: with epilogue + producer 10 warps, i.e. <= 3 per SM sub-partition -> 168 registers
constexpr int MK_CTHREADS = MK_CW * 32;
constexpr int MK_STHREADS = MK_CTHREADS + 32; // + 1 epilogue warp: the threads that take part in phase-level barriers
constexpr int MK_THREADS = MK_STHREADS + 64;  // + 1 producer warp + 1 L2 prefetch warp: 11 warps, <= 3 per SM sub-partition
constexpr int MK_NRED = 4;                    // reduction buffers in flight between the consumer warps and the epilogue warp
constexpr int MK_MAXSLOTS = 24;
constexpr int MK_GROUP = 8;                   // weight rows per stage (= n of the mma)
constexpr int MK_MAXGROUPS = 8;               // row groups of one CTA in the small-matrix phases (bias / residual preload)
constexpr int MK_DEFAULT_PF = 16;               // defaults of PADT_MEGA_PF / PADT_MEGA_FLAGS (see MegaArgs)
constexpr int MK_DEFAULT_FLAGS = 1 | 4 | 8 | 64 | 128; // MF_EVICT_FIRST | MF_KV_SMEM | MF_PF_LINES | MF_STAGE_MIN | MF_SPLIT
constexpr int MK_DEFAULT_PF_PACE = 0;
constexpr int MK_ATTN_SMEM = 2 * DA_TILE * DA_LD * 2 + DA_MAXG * DA_HD * 4 + DA_MAXG * DA_TILE * 4;

struct StageDesc;
struct MegaArgs {
    DecodeParams p;
    const DecodeLayerW* layers;   // device copy
    unsigned int* gbar;           // grid barrier counter, zeroed by a memset node before every launch
    int slot_bytes, n_slots, x_bytes;
    const struct StageDesc* stage_tab;   // [sum stages] descriptors, CTA c owns [stage_off[c], stage_off[c+1])
    const int* stage_off;
    int dbg;                      // PADT_MEGA_TIMELINE bits: 1 print cycle breakdown, 2 no epilogue / mma, 4 no grid barrier, 8 no attention, 64 no mma
    int pf_depth;                 // L2 prefetch warp: stages kept in flight towards L2 beyond the ring (0 = off)
    int flags;                    // MF_* bits
    int pf_pace;                  // minimum cycles between two stage prefetches (keeps the request queues short during a stall)
    int attn_off;                 // byte offset of the attention scratch inside the activation region
    int hraw_off;                 // byte offset (from the activation region) of the raw residual rows [B][D] bf16
    // LL exchange buffers (uint64 words), two parity copies `ll_stride` words apart (layer parity): word offsets of the vectors
    unsigned long long* ll;
    unsigned int ll_stride, ll_hA, ll_hB, ll_qkv, ll_attn, ll_act, ll_part;
    // every broadcast vector (read by ALL CTAs) is published in `ll_nrep` replicas `ll_rep` words apart; CTA c reads replica c % ll_nrep.
    // 148 CTAs polling the same 64 lines serialise in the L2 slices that own them (~1 us per round of the 8 KB residual row, and the
    // pollers delay the very stores they wait for); with 16 replicas a line has ~9 readers.
    unsigned int ll_rep, ll_nrep;
};
enum { MF_EVICT_FIRST = 1, MF_NO_AREG = 2, MF_KV_SMEM = 4, MF_PF_LINES = 8, MF_PF_EVICT_LAST = 16, MF_STAGE_MIN = 64, MF_SPLIT = 128 };

using CSync = NamedSync<1, MK_STHREADS>;   // consumer warps + epilogue warp (start / end of the launch only)
using KSync = NamedSync<2, MK_CTHREADS>;   // consumer warps only: the epilogue warp runs behind them, coupled only through the reduction mbarriers


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
// -DPADT_MK_WATCHDOG: every wait site reports itself and traps after ~2 s instead of hanging the GPU (debug builds only)
#ifdef PADT_MK_WATCHDOG
__device__ int g_wd_phase[160];
__device__ int g_wd_sub[160];     // finer position inside the phase (written by tid 0)
__device__ int g_wd_epi[160];     // groups finished by the epilogue warp
__device__ int g_wd_prod[160];    // last stage issued by producer lane 0
#define WD_MARK(arr, v) do { ((volatile int*)arr)[blockIdx.x] = (v); } while (0)
#define WD_DECL(site) long long _wd_t0 = clock64(); const int _wd_site = (site)
#define WD_CHECK()                                                                                                              \
    do {                                                                                                                        \
        if (clock64() - _wd_t0 > 4000000000LL) {                                                                                \
            printf("[watchdog] site %d cta %d warp %d lane %d phase %d\n", _wd_site, blockIdx.x, threadIdx.x >> 5, threadIdx.x & 31, \
                   g_wd_phase[blockIdx.x]);                                                                                     \
            if (blockIdx.x == 0 || blockIdx.x == 1) {                                                                           \
                for (int _c = 0; _c < (int)gridDim.x; ++_c) printf("[watchdog] cta %d phase %d sub %d epi_groups %d producer_stage %d\n", _c, ((volatile int*)g_wd_phase)[_c], ((volatile int*)g_wd_sub)[_c], ((volatile int*)g_wd_epi)[_c], ((volatile int*)g_wd_prod)[_c]); \
                long long _t1 = clock64();                                                                                      \
                while (clock64() - _t1 < 200000000LL) {}                                                                        \
            } else {                                                                                                            \
                long long _t1 = clock64();                                                                                      \
                while (clock64() - _t1 < 400000000LL) {}                                                                        \
            }                                                                                                                   \
            __trap();                                                                                                           \
        }                                                                                                                       \
    } while (0)
__device__ __forceinline__ void mbar_wait_wd(uint32_t bar, uint32_t parity, int site) {
    WD_DECL(site);
    for (;;) {
        uint32_t ok;
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) break;
        WD_CHECK();
    }
}
#else
#define WD_MARK(arr, v) do { } while (0)
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity, int site = 0) {
#if defined(PADT_MK_WATCHDOG) && PADT_MK_WATCHDOG >= 2
    mbar_wait_wd(smem_u32(b), parity, site);
    return;
#endif
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
__device__ __forceinline__ void tma_bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
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
    int row_lo;         // first global output row of this CTA's slice (group-aligned partition)
    int rows;           // valid rows in the slice (the last group may be ragged)
    int groups;         // ceil(rows / 8)
    int K, Kseg, nseg;  // row length, columns per stage, stages per row group
    // stage-granular view: this CTA consumes stages [st_lo, st_hi) of the matrix, stage = group * nseg + K segment.  With the
    // balanced partition (MF_SPLIT) the range may begin / end inside a row group; that group is then shared with a neighbour CTA.
    int st_lo, st_hi;
    int row_begin, n_rows;   // first output row of the matrix, rows of the matrix
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
                                         int slot_bytes, bool split = false) {
    Seg s;
    const unsigned int groups = (unsigned int)(n_rows + MK_GROUP - 1) / MK_GROUP;       // groups * nseg * G < 2^31 for every matrix here
    const int lo = (int)((groups * (unsigned int)cta) / (unsigned int)G), hi = (int)((groups * (unsigned int)(cta + 1)) / (unsigned int)G);
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
    s.row_begin = row_begin; s.n_rows = n_rows;
    if (split && img && groups >= (unsigned int)G) {     // every CTA gets >= nseg stages, so a row group spans at most two CTAs
        const unsigned int S = groups * (unsigned int)s.nseg;
        s.st_lo = (int)((S * (unsigned int)cta) / (unsigned int)G);
        s.st_hi = (int)((S * (unsigned int)(cta + 1)) / (unsigned int)G);
    } else {
        s.st_lo = lo * s.nseg; s.st_hi = hi * s.nseg;
    }
    return s;
}

enum { SEG_QKV = 0, SEG_O = 1, SEG_GU = 2, SEG_DOWN = 3 };

__device__ __forceinline__ Seg layer_seg(const MegaArgs& a, const DecodeLayerW& w, int which, int cta, int G) {
    const DecodeParams& p = a.p;
    // Balanced (stage-granular) partition only for down_proj: it is the HBM-bound phase with the coarsest quantisation (256 row groups
    // of 8 stages for 148 CTAs).  Measured on q/k/v, o_proj and gate/up (2 stages per group, one per half-CTA) it LOSES: a shared
    // group with a single stage keeps one half-CTA idle, so the phase gets longer by one stage time.
    const bool split = (a.flags & MF_SPLIT) != 0;
    switch (which) {
        case SEG_QKV: return make_seg(w.wqkv, w.iqkv, p.nseg_d, 0, (p.nh + 2 * p.nkv) * p.hd, p.D, cta, G, a.slot_bytes, false);
        case SEG_O: return make_seg(w.wo, w.io, p.nseg_a, 0, p.D, p.nh * p.hd, cta, G, a.slot_bytes, false);
        case SEG_GU: return make_seg(w.wgu, w.igu, p.nseg_d, 0, 2 * p.I, p.D, cta, G, a.slot_bytes, false);
        default: return make_seg(w.wdown, w.idown, p.nseg_i, 0, p.D, p.I, cta, G, a.slot_bytes, split);
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

__device__ void producer_loop(const MegaArgs& a, uint8_t* ring, uint64_t* full, uint64_t* empty, int cta, int G, int lane,
                              volatile unsigned int* prog, long long* t_wait_empty) {
    const StageDesc* tab = a.stage_tab + a.stage_off[cta];
    const int n = a.stage_off[cta + 1] - a.stage_off[cta];
    const int ns = a.n_slots;
    if (lane < ns) {
        uint8_t* dst = ring + (size_t)lane * a.slot_bytes;
        uint32_t parity = 0;
        const bool hint = (a.flags & MF_EVICT_FIRST) != 0;
        const uint64_t pol = l2_policy_evict_first();
        long long tw = 0;
        for (int i = lane; i < n; i += ns, parity ^= 1) {
            StageDesc d = tab[i];
            long long tq = (a.dbg & 1) ? clock64() : 0;
            mbar_wait(&empty[lane], parity ^ 1, 1);
            if (a.dbg & 1) tw += clock64() - tq;
            mbar_expect_tx(&full[lane], d.bytes);
            if (hint) tma_bulk_g2s_hint(dst, d.src, d.bytes, &full[lane], pol);
            else tma_bulk_g2s(dst, d.src, d.bytes, &full[lane]);
            if (lane == 0) { *prog = (unsigned int)i; WD_MARK(g_wd_prod, i); }     // issue frontier, read by the L2 prefetch warp
        }
        if (lane == 0) { *prog = (unsigned int)n; *t_wait_empty = tw; }
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
            mbar_wait(&empty[slot], par ^ 1, 2);
            mbar_expect_tx(&full[slot], seg_bytes * rows);
            uint8_t* dst = ring + (size_t)slot * a.slot_bytes;
            const bf16* src = s.base + (size_t)(g * MK_GROUP) * s.K + (size_t)sg * s.Kseg;
            for (int r = 0; r < rows; ++r) tma_bulk_g2s(dst + r * row_stride, src + (size_t)r * s.K, seg_bytes, &full[slot]);
        }
    }
}

// ---------------------------------------------------------------- L2 prefetch warp
// The ring holds ~3.6 us of HBM stream; a phase boundary (grid barrier + activation staging, or the attention phase) stalls the
// consumers for longer than that, and a full ring stops the producer, so HBM would sit idle.  This warp keeps the stream
// going during such stalls: it walks the same descriptor table `pf_depth` stages ahead of the producer's issue frontier and
// asks the L2 (126 MB) to fetch those stages; when the consumers resume, the ring refills at L2 rather than HBM rate.
__device__ void prefetch_loop(const MegaArgs& a, int cta, int lane, volatile unsigned int* prog) {
    const StageDesc* tab = a.stage_tab + a.stage_off[cta];
    const int n = a.stage_off[cta + 1] - a.stage_off[cta];
    const int depth = a.pf_depth;
    if (depth <= 0) return;
    const bool keep = (a.flags & MF_PF_EVICT_LAST) != 0;
    int pf = a.n_slots;                          // the first ring fill is issued by the producer itself
    while (pf < n) {
        int lim = (int)*prog + a.n_slots + depth;
        lim = lim > n ? n : lim;
        if (pf >= lim) { __nanosleep(64); continue; }
        if (a.pf_pace > 0) lim = pf + 1;        // paced: one stage at a time
        lim = lim > pf + 32 ? pf + 32 : lim;    // one descriptor per lane per round (a single coalesced load, no serial latency chain)
        StageDesc mine = {nullptr, 0, 0};
        if (pf + lane < lim) mine = tab[pf + lane];
        if (a.flags & MF_PF_LINES) {
            // through the LSU: one 128-byte line per lane per instruction; keeps the TMA queue free for the ring copies
            for (int i = 0; i < lim - pf; ++i) {
                const char* src = reinterpret_cast<const char*>(__shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)mine.src, i));
                const uint32_t bytes = __shfl_sync(0xffffffffu, mine.bytes, i);
                if (keep) {
                    for (uint32_t o = (uint32_t)lane * 128; o < bytes; o += 32 * 128)
                        asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(src + o) : "memory");
                } else {
                    for (uint32_t o = (uint32_t)lane * 128; o < bytes; o += 32 * 128)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(src + o) : "memory");
                }
            }
        } else if (pf + lane < lim) {
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(mine.src), "r"(mine.bytes) : "memory");
        }
        pf = lim;
        if (a.pf_pace > 0) {
            long long t = clock64();
            while (clock64() - t < a.pf_pace) __nanosleep(32);
        }
    }
}

// ---------------------------------------------------------------- consumer helpers
// Shared-memory operations on 32-bit shared-window addresses: the addresses are computed once per phase, so the inner loops
// carry no generic->shared conversions or index arithmetic (the v3 loop spent ~35 instructions per two k-steps on them and was
// ISSUE bound at ~700 cycles per 16 KB stage, i.e. no faster than HBM delivers a stage).
__device__ __forceinline__ void mbar_wait_s(uint32_t bar, uint32_t parity, int site = 0) {
#if defined(PADT_MK_WATCHDOG) && PADT_MK_WATCHDOG >= 2
    mbar_wait_wd(bar, parity, site);
    return;
#endif
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_s(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void ldsm_x4_s(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_s(uint32_t& r0, uint32_t& r1, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];\n" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ uint32_t lds32_s(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
// D += A * B with A = {a0, 0, a2, 0}: only the first 8 rows of the 16-row A tile (the batch rows) are populated
__device__ __forceinline__ void mma_rows8(float* c, uint32_t a0, uint32_t a2, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%5}, {%7,%8}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(0u), "r"(a2), "r"(b0), "r"(b1));
}
// ---------------------------------------------------------------- LL words: {epoch, payload} in one 8-byte single-copy-atomic word
// Producer: st.relaxed.gpu of the word (no fence, no flag).  Consumer: ld.relaxed.gpu until the epoch matches.  Epochs are
// unique per (decode launch, phase) and never 0 (the buffers are zeroed at prefill), so a stale word can never be accepted.
// A poll that spins for ~10 s traps instead of hanging the GPU.
typedef unsigned long long ull;
constexpr unsigned int LL_SPIN_LIMIT = 1u << 25;
__device__ int g_ll_sleep = 0;    // PADT_MEGA_LL_SLEEP: nanoseconds a poller sleeps after a failed poll (experiment knob)
__device__ __forceinline__ void ll_backoff() {
    const int ns = g_ll_sleep;
    if (ns > 0) __nanosleep(ns);
}
__device__ __noinline__ void ll_timeout(const void* p, unsigned int epoch, int site) {
    printf("[decode_mega] LL poll timed out: cta %d thread %d site %d epoch %u addr %p\n", blockIdx.x, threadIdx.x, site, epoch, p);
    __trap();
}
__device__ __forceinline__ void ll_store(ull* p, unsigned int epoch, uint32_t payload) {
    const ull w = ((ull)epoch << 32) | (ull)payload;
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ void ll_store2(ull* p, unsigned int epoch, uint32_t a, uint32_t b) {
    const ull w0 = ((ull)epoch << 32) | (ull)a, w1 = ((ull)epoch << 32) | (ull)b;
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1,%2};" ::"l"(p), "l"(w0), "l"(w1) : "memory");
}
__device__ __forceinline__ void ll_ld2(const ull* p, ull& x, ull& y) {
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(x), "=l"(y) : "l"(p) : "memory");
}
__device__ __forceinline__ bool ll_ok(ull x, unsigned int epoch) { return (unsigned int)(x >> 32) == epoch; }
// two words (16 bytes)
__device__ __forceinline__ void ll_wait2(const ull* p, unsigned int epoch, uint32_t& a, uint32_t& b, int site) {
    ull x, y;
    unsigned int spins = 0;
    for (;;) {
        ll_ld2(p, x, y);
        if (ll_ok(x, epoch) && ll_ok(y, epoch)) break;
        ll_backoff(); if (++spins > LL_SPIN_LIMIT) ll_timeout(p, epoch, site);
    }
    a = (uint32_t)x; b = (uint32_t)y;
}
// four consecutive words (32 bytes = 8 bf16 payload values), both 16-byte loads in flight together
__device__ __forceinline__ uint4 ll_wait4(const ull* p, unsigned int epoch, int site) {
    ull x0, y0, x1, y1;
    unsigned int spins = 0;
    ll_ld2(p, x0, y0);
    ll_ld2(p + 2, x1, y1);
    for (;;) {
        const bool ok0 = ll_ok(x0, epoch) && ll_ok(y0, epoch), ok1 = ll_ok(x1, epoch) && ll_ok(y1, epoch);
        if (ok0 && ok1) break;
        if (!ok0) ll_ld2(p, x0, y0);
        if (!ok1) ll_ld2(p + 2, x1, y1);
        ll_backoff(); if (++spins > LL_SPIN_LIMIT) ll_timeout(p, epoch, site);
    }
    return make_uint4((uint32_t)x0, (uint32_t)y0, (uint32_t)x1, (uint32_t)y1);
}
// up to three 8-value chunks per thread with all six 16-byte loads in flight (keeps a long vector -- the SwiGLU activations --
// at ONE L2 round trip per thread instead of one per chunk)
__device__ __forceinline__ void ll_wait4x3(const ull* p0, const ull* p1, const ull* p2, int n, unsigned int epoch, uint4* out, int site) {
    ull r[12];
    unsigned int spins = 0;
    const ull* ps[3] = {p0, p1, p2};
#pragma unroll
    for (int i = 0; i < 3; ++i)
        if (i < n) { ll_ld2(ps[i], r[4 * i], r[4 * i + 1]); ll_ld2(ps[i] + 2, r[4 * i + 2], r[4 * i + 3]); }
    for (;;) {
        bool all = true;
#pragma unroll
        for (int i = 0; i < 3; ++i)
            if (i < n) {
                const bool ok0 = ll_ok(r[4 * i], epoch) && ll_ok(r[4 * i + 1], epoch), ok1 = ll_ok(r[4 * i + 2], epoch) && ll_ok(r[4 * i + 3], epoch);
                if (!ok0) ll_ld2(ps[i], r[4 * i], r[4 * i + 1]);
                if (!ok1) ll_ld2(ps[i] + 2, r[4 * i + 2], r[4 * i + 3]);
                all = all && ok0 && ok1;
            }
        if (all) break;
        ll_backoff(); if (++spins > LL_SPIN_LIMIT) ll_timeout(p0, epoch, site);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i)
        if (i < n) out[i] = make_uint4((uint32_t)r[4 * i], (uint32_t)r[4 * i + 1], (uint32_t)r[4 * i + 2], (uint32_t)r[4 * i + 3]);
}

// activation rows -> shared memory, row stride K + 8 elements (the 16-byte pad keeps the A-fragment loads conflict free).
// Source of row b: LL words `ll + b * K/2` (polled until their epoch is `epoch`) or, when ll == nullptr, the plain bf16 row rows[b]
// (layer 0: the embedding / prototype row of the sampled token).
// ln_w != null: RMSNorm prologue, bit-identical to rmsnorm_kernel / the decode.cu prologue (rms_apply8; per-warp partial sums of
// squares -> shared memory -> fixed-order total).  Every element of the row is fetched by exactly ONE thread of the CTA
// (K <= 16 * MK_CTHREADS, checked by the plan); the raw row is kept in `hraw` as the residual source of the epilogues.
__device__ __forceinline__ void stage_norm(const bf16* const* rows, const ull* ll, unsigned int epoch, const bf16* ln_w, float eps, int K, int B,
                                           bf16* xs, bf16* hraw, int tid, int lane, bf16* copy0, bf16* copy1, float* s_part) {
    if (tid >= MK_CTHREADS) return;
    const int x_ld = K + 8;
    const int nch = K >> 3, warp = tid >> 5;
    const int c0 = tid, c1 = tid + MK_CTHREADS;               // K <= 16 * MK_CTHREADS (checked by the plan): at most two chunks per thread
    const bool m0 = c0 < nch, m1 = c1 < nch;
    uint4 wv0 = make_uint4(0, 0, 0, 0), wv1 = wv0;
    if (m0) wv0 = *reinterpret_cast<const uint4*>(ln_w + c0 * 8);
    if (m1) wv1 = *reinterpret_cast<const uint4*>(ln_w + c1 * 8);
    for (int b = 0; b < B; ++b) {
        uint4 xv0 = make_uint4(0, 0, 0, 0), xv1 = xv0;
        if (ll) {
            const ull* row = ll + (size_t)b * (K >> 1);
            uint4 v[3];
            if (m0) {
                ll_wait4x3(row + (size_t)c0 * 4, row + (size_t)c1 * 4, row, m1 ? 2 : 1, epoch, v, 10);
                xv0 = v[0];
                if (m1) xv1 = v[1];
            }
        } else {
            if (m0) xv0 = *reinterpret_cast<const uint4*>(rows[b] + c0 * 8);
            if (m1) xv1 = *reinterpret_cast<const uint4*>(rows[b] + c1 * 8);
        }
        float ss;
        {
            float2 e0 = bf2x_to_f2(xv0.x), e1 = bf2x_to_f2(xv0.y), e2 = bf2x_to_f2(xv0.z), e3 = bf2x_to_f2(xv0.w);
            ss = e0.x * e0.x + e0.y * e0.y + e1.x * e1.x + e1.y * e1.y + e2.x * e2.x + e2.y * e2.y + e3.x * e3.x + e3.y * e3.y;
            e0 = bf2x_to_f2(xv1.x); e1 = bf2x_to_f2(xv1.y); e2 = bf2x_to_f2(xv1.z); e3 = bf2x_to_f2(xv1.w);
            ss += e0.x * e0.x + e0.y * e0.y + e1.x * e1.x + e1.y * e1.y + e2.x * e2.x + e2.y * e2.y + e3.x * e3.x + e3.y * e3.y;
        }
        ss = warp_sum(ss);
        if (lane == 0) s_part[warp] = ss;
        asm volatile("bar.sync 2, %0;" ::"n"(MK_CTHREADS) : "memory");
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < MK_CW; ++w) tot += s_part[w];
        const float rstd = 1.0f / sqrtf(tot / (float)K + eps);
        if (m0) {
            uint4 o = rms_apply8(xv0, wv0, rstd);
            *reinterpret_cast<uint4*>(xs + (size_t)b * x_ld + c0 * 8) = o;
            if (hraw) *reinterpret_cast<uint4*>(hraw + (size_t)b * K + c0 * 8) = xv0;
            if (copy0) *reinterpret_cast<uint4*>(copy0 + (size_t)b * K + c0 * 8) = o;
            if (copy1) *reinterpret_cast<uint4*>(copy1 + (size_t)b * K + c0 * 8) = o;
        }
        if (m1) {
            uint4 o = rms_apply8(xv1, wv1, rstd);
            *reinterpret_cast<uint4*>(xs + (size_t)b * x_ld + c1 * 8) = o;
            if (hraw) *reinterpret_cast<uint4*>(hraw + (size_t)b * K + c1 * 8) = xv1;
            if (copy0) *reinterpret_cast<uint4*>(copy0 + (size_t)b * K + c1 * 8) = o;
            if (copy1) *reinterpret_cast<uint4*>(copy1 + (size_t)b * K + c1 * 8) = o;
        }
        if (b + 1 < B) asm volatile("bar.sync 2, %0;" ::"n"(MK_CTHREADS) : "memory");   // s_part is reused by the next row
    }
}
// plain copy of LL rows (attention output, SwiGLU activations): up to three 8-value chunks per thread per round trip
__device__ __forceinline__ void stage_copy(const ull* ll, unsigned int epoch, int K, int B, bf16* xs, int tid) {
    if (tid >= MK_CTHREADS) return;
    const int x_ld = K + 8, nch = K >> 3;
    for (int b = 0; b < B; ++b) {
        const ull* row = ll + (size_t)b * (K >> 1);
        for (int c0 = tid; c0 < nch; c0 += 3 * MK_CTHREADS) {
            const int c1 = c0 + MK_CTHREADS, c2 = c0 + 2 * MK_CTHREADS;
            const int n = 1 + (c1 < nch) + (c2 < nch);
            uint4 v[3];
            ll_wait4x3(row + (size_t)c0 * 4, row + (size_t)c1 * 4, row + (size_t)c2 * 4, n, epoch, v, 11);
            *reinterpret_cast<uint4*>(xs + (size_t)b * x_ld + c0 * 8) = v[0];
            if (n > 1) *reinterpret_cast<uint4*>(xs + (size_t)b * x_ld + c1 * 8) = v[1];
            if (n > 2) *reinterpret_cast<uint4*>(xs + (size_t)b * x_ld + c2 * 8) = v[2];
        }
    }
}

enum { EP_BIAS = 0, EP_RESID = 1, EP_SWIGLU = 2, EP_LOGITS = 3 };

struct SmemMap {                 // 32-bit shared-window addresses / pointers of the fixed regions
    uint32_t ring, full, empty, red_full, red_empty, xs;
    float* red;                  // [MK_NRED][MK_CW][B][8] partial sums (K slices of the 16 consumer warps)
    float* pre;                  // [8 batch][MK_MAXGROUPS*8] preloaded bias / residual values (epilogue warp only)
};

struct ArgMax { float v[2][32]; int i[2][32]; int lo[2][32], hi[2][32]; };   // epilogue-warp state in shared memory: [pass][lane] -> batch row pass*4 + lane/8

// Consume one segment from the ring; `it` / `grp` are CTA-wide stage / row-group counters every thread advances identically.
// Consumer warps (0..15): two half-CTAs of 8 warps alternate over the stages (even / odd K segment); inside a stage every warp
// multiplies its own k16-steps (ldmatrix + mma.sync, two independent accumulator chains).  A warp always owns the SAME K slice
// of a segment, so when the segment has <= 2 K segments and 64 k-steps (K = 2048: q/k/v, o_proj, gate/up, LM head) its
// activation fragments are loaded into 16 registers ONCE per phase and the per-stage loop is 4 x (ldmatrix.x4 + 2 mma).
// After the last stage of a row group a warp drops its 8 partial sums into a reduction buffer and signals an mbarrier -- it
// never waits for the other warps.  Epilogue warp (16): waits for the 16 partials of a group, adds them in warp order and
// applies the HF epilogue.
struct ConsumeArgs {   // passed BY VALUE to the noinline consume_seg: isolates the register allocation of the hot loop from the phase loop
    Seg s;
    SmemMap sm;
    ArgMax* am;
    const bf16* bias; ull* ll_out; const bf16* hraw;
    unsigned long long* xpart; unsigned int* gbar;
    uint32_t it, grp;
    int ep, tid, n_out, B, vocab, n_slots, slot_bytes, flags, dbg;
    unsigned int epoch, wait_layers, ll_rep, ll_nrep;
    volatile unsigned int* epi_done;   // shared memory: row groups the epilogue warp has finished (read by the staging code before it overwrites hraw)
};
__device__ __forceinline__ void consume_seg(const ConsumeArgs ca) {
    const Seg& s = ca.s;
    const SmemMap& sm = ca.sm;
    ArgMax& am = *ca.am;
    const int ep = ca.ep, tid = ca.tid, n_out = ca.n_out;
    const bf16* bias = ca.bias; ull* ll_out = ca.ll_out; const bf16* hraw = ca.hraw;
    const unsigned int epoch = ca.epoch, wait_layers = ca.wait_layers;
    uint32_t it = ca.it, grp = ca.grp;
    const int warp = tid >> 5, lane = tid & 31;
    const int B = ca.B;
    const int n_slots = ca.n_slots;
    uint32_t slot = it % (uint32_t)n_slots, parity = (it / (uint32_t)n_slots) & 1u;
    const int n_st = s.st_hi - s.st_lo;
    const int g_first = s.st_lo / s.nseg;                                   // first (possibly shared) row group of this CTA
    const int n_grp = n_st > 0 ? (s.st_hi - 1) / s.nseg - g_first + 1 : 0;  // row groups touched (whole or partial)
    it += (uint32_t)n_st;
    uint32_t g_idx = grp;
    grp += (uint32_t)n_grp;
    if (warp == MK_CW) {
        // ================= epilogue warp: lane -> (batch row eb = pass*4 + lane/8, weight row en = lane%8)
        // A row group whose K segments are shared with a neighbour CTA (balanced partition): the CTA holding the TAIL segments
        // publishes its fp32 partial sums (it meets that group first), the CTA holding the HEAD segments adds them to its own
        // (it meets the group last, the partial has long been published) and applies the epilogue.  Fixed order: head + tail.
        const int en = lane & 7;
        const bool tail_first = n_grp > 0 && (s.st_lo % s.nseg) != 0;      // group 0 of this CTA: its head lives in CTA - 1
        const bool head_last = n_grp > 0 && (s.st_hi % s.nseg) != 0;       // last group: its tail lives in CTA + 1
        if (wait_layers > 0) {
            // WAR guard of the LL buffers (two parity copies): before this CTA writes anything of layer l, every CTA must have
            // finished reading layer l - 2, i.e. passed the q/k/v staging of layer l - 1.  The CTAs move in lock step through
            // their data dependencies, so this is true long before it is checked (one acquire load, overlapped with the first mma
            // group); it only ever waits when a CTA without rows of its own has fallen a full layer behind.
            if (lane == 0) {
                unsigned int spins = 0;
                while (ld_acquire_u32(ca.gbar) < wait_layers * gridDim.x)
                    ll_backoff(); if (++spins > LL_SPIN_LIMIT) ll_timeout(ca.gbar, wait_layers, 12);
            }
            __syncwarp();
        }
        if (ep == EP_BIAS) {   // preload so that the per-group epilogue never waits on global memory
            for (int pass = 0; pass * 4 < B; ++pass) {
                int eb = pass * 4 + (lane >> 3);
                if (eb < B)
                    for (int gi = tail_first ? 1 : 0; gi < n_grp && gi < MK_MAXGROUPS; ++gi) {
                        int r = (g_first + gi) * 8 + en;
                        if (r < s.n_rows) sm.pre[eb * (MK_MAXGROUPS * 8) + gi * 8 + en] = bias ? bf2f(bias[s.row_begin + r]) : 0.f;
                    }
            }
            __syncwarp();
        }
        const int out_words = n_out >> 1;                                   // LL words per batch row of this phase's output vector
        for (int gi = 0; gi < n_grp; ++gi, ++g_idx) {
            const uint32_t buf = g_idx % MK_NRED;
            const bool is_tail = gi == 0 && tail_first, is_head = gi == n_grp - 1 && head_last;
            float nb0[2] = {0.f, 0.f}, nb1[2] = {0.f, 0.f};                 // neighbour's partial: rows en / en + 1, batch rows of pass 0 / 1
            if (is_head) {
                // fetched while the consumer warps are still multiplying this group.  Each partial travels as ONE 64-bit word
                // {epoch, fp32 bits}: no flag, no fence on the writer, one L2 round trip on the reader.
#pragma unroll
                for (int pass = 0; pass < 2; ++pass) {
                    const int eb = pass * 4 + (lane >> 3);
                    if (eb < B) {
                        const unsigned long long* xp = ca.xpart + ((size_t)(blockIdx.x + 1) * 8 + eb) * 8;
                        unsigned long long w0, w1;
                        unsigned int spins = 0;
                        do {
                            asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w0) : "l"(xp + en) : "memory");
                            asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w1) : "l"(xp + (en | 1)) : "memory");
                            ll_backoff(); if (++spins > LL_SPIN_LIMIT) ll_timeout(xp, epoch, 13);
                        } while ((unsigned int)(w0 >> 32) != epoch || (unsigned int)(w1 >> 32) != epoch);
                        nb0[pass] = __uint_as_float((unsigned int)w0);
                        nb1[pass] = __uint_as_float((unsigned int)w1);
                    }
                }
            }
            mbar_wait_s(sm.red_full + buf * 8, (g_idx / MK_NRED) & 1, 3);
            const float* red = sm.red + (size_t)buf * MK_CW * B * 8;
            const int r = (g_first + gi) * 8 + en;
            const int n = s.row_begin + r;
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {              // static indices keep the arg-max state in registers
                if (pass * 4 >= B) break;                       // warp-uniform
                const int eb = pass * 4 + (lane >> 3);
                const bool active = eb < B;
                float sum = 0.f;
                if (active) {
#pragma unroll
                    for (int w = 0; w < MK_CW; ++w) sum += red[(w * B + eb) * 8 + en];
                }
                if (is_tail) {                                  // publish: the neighbour finishes this group (warp-uniform branch)
                    if (active) {
                        const unsigned long long w = ((unsigned long long)epoch << 32) | (unsigned long long)__float_as_uint(sum);
                        asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(ca.xpart + ((size_t)blockIdx.x * 8 + eb) * 8 + en), "l"(w) : "memory");
                    }
                    continue;
                }
                const bool row_ok = active && r < s.n_rows;
                sum += nb0[pass];
                if (ep == EP_LOGITS) {   // bf16-rounded logit, -inf outside the row's own VRT range, first max wins
                    float l0 = rbf(sum);
                    bool ok = row_ok && (n < ca.vocab || (n - ca.vocab >= am.lo[pass][lane] && n - ca.vocab < am.hi[pass][lane]));
                    const float bv = am.v[pass][lane];
                    if (ok && (l0 > bv || (l0 == bv && n < am.i[pass][lane]))) { am.v[pass][lane] = l0; am.i[pass][lane] = n; }
                    continue;                                    // warp-uniform (ep)
                }
                // the value of this lane's output row as bf16 bits; two neighbouring rows travel in one LL word
                uint32_t val = 0;
                if (ep == EP_BIAS) {
                    if (row_ok) val = (uint32_t)__bfloat16_as_ushort(f2bf(sum + sm.pre[eb * (MK_MAXGROUPS * 8) + gi * 8 + en]));
                } else if (ep == EP_RESID) {
                    if (row_ok) val = (uint32_t)__bfloat16_as_ushort(f2bf(bf2f(hraw[(size_t)eb * n_out + n]) + rbf(sum)));
                } else {   // EP_SWIGLU: even row = gate_j, odd row = up_j of the interleaved weight
                    if (row_ok && (en & 1) == 0) {
                        float up = nb1[pass];
#pragma unroll
                        for (int w = 0; w < MK_CW; ++w) up += red[(w * B + eb) * 8 + en + 1];
                        float gv = rbf(sum), uv = rbf(up);
                        val = (uint32_t)__bfloat16_as_ushort(f2bf(rbf(silu_f(gv)) * uv));
                    }
                }
                const int dist = ep == EP_SWIGLU ? 2 : 1;                     // lane distance of the value that shares the word
                const uint32_t hi = __shfl_down_sync(0xffffffffu, val, dist);
                const int oi = ep == EP_SWIGLU ? (n >> 1) : n;                // index of this lane's value in the output vector
                if (row_ok && (en & (2 * dist - 1)) == 0) {
                    ull* dst = ll_out + (size_t)eb * out_words + (oi >> 1);
                    for (unsigned int r = 0; r < ca.ll_nrep; ++r, dst += ca.ll_rep) ll_store(dst, epoch, val | (hi << 16));
                }
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive_s(sm.red_empty + buf * 8);
                *ca.epi_done = g_idx + 1;
                WD_MARK(g_wd_epi, (int)g_idx);
            }
        }
        return;
    }
    // ================= consumer warps
    const int g = lane >> 2, t = lane & 3;
    const uint32_t row_stride = (uint32_t)s.Kseg * 2 + 16;
    const int wi = warp;                                                    // every warp takes part in every stage and owns one K slice of it
    const int ksteps = s.Kseg >> 4;
    const int k0 = (ksteps * wi) / MK_CW, k1 = (ksteps * (wi + 1)) / MK_CW;
    const bool live = g < B;
    // per-lane byte offset of the ldmatrix.x4 rows inside a stage: matrices 0/1 = k-step kk (k 0-7 / 8-15), 2/3 = k-step kk+1
    const uint32_t w_off = (uint32_t)(lane & 7) * row_stride + (uint32_t)(((lane >> 4) * 16 + ((lane >> 3) & 1) * 8) * 2) + (uint32_t)k0 * 32u;
    const uint32_t x_row = sm.xs + (uint32_t)((g * (s.K + 8) + 2 * t + k0 * 16) * 2);
    // 64 k-steps per K segment (K = 2048 phases at 3B: q/k/v, o_proj, gate/up, LM head; 71 % of the bytes): 8 k-steps per warp, whose 16
    // activation-fragment registers are (re)loaded from shared memory BEFORE the wait for the stage, so the loop after the wait is
    // 4 x (ldmatrix.x4 + 2 mma) with no address arithmetic
    const bool reg_path = ksteps == 8 * MK_CW && !(ca.flags & MF_NO_AREG);
    const bool skip_math = (ca.dbg & (2 | 64)) != 0;
    int sg = s.st_lo % s.nseg;                                              // the first group may start in the middle (balanced partition)
    for (int idx = s.st_lo; idx < s.st_hi; ++g_idx) {
        float c[4] = {0.f, 0.f, 0.f, 0.f}, c2[4] = {0.f, 0.f, 0.f, 0.f};
        for (; sg < s.nseg && idx < s.st_hi; ++sg, ++idx) {
            const uint32_t my_slot = slot, my_parity = parity;
            if (++slot == (uint32_t)n_slots) { slot = 0; parity ^= 1u; }
            uint32_t areg[16];
            if (reg_path) {
                const uint32_t xb = x_row + (uint32_t)(sg * s.Kseg * 2);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    areg[2 * i] = live ? lds32_s(xb + i * 32) : 0u;
                    areg[2 * i + 1] = live ? lds32_s(xb + i * 32 + 16) : 0u;
                }
            }
            mbar_wait_s(sm.full + my_slot * 8, my_parity, 4);
            if (!skip_math) {
                const uint32_t wb = sm.ring + my_slot * (uint32_t)ca.slot_bytes + w_off;
                if (reg_path) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint32_t b0, b1, b2, b3;
                        ldsm_x4_s(b0, b1, b2, b3, wb + j * 64);
                        mma_rows8(c, areg[4 * j], areg[4 * j + 1], b0, b1);      // two independent accumulator chains
                        mma_rows8(c2, areg[4 * j + 2], areg[4 * j + 3], b2, b3);
                    }
                } else {
                    uint32_t wa = wb, xa = x_row + (uint32_t)(sg * s.Kseg * 2);
                    int kk = k0;
                    for (; kk + 1 < k1; kk += 2, wa += 64, xa += 64) {
                        uint32_t b0, b1, b2, b3;
                        ldsm_x4_s(b0, b1, b2, b3, wa);
                        uint32_t a0 = 0, a2 = 0, a4 = 0, a6 = 0;
                        if (live) { a0 = lds32_s(xa); a2 = lds32_s(xa + 16); a4 = lds32_s(xa + 32); a6 = lds32_s(xa + 48); }
                        mma_rows8(c, a0, a2, b0, b1);
                        mma_rows8(c2, a4, a6, b2, b3);
                    }
                    if (kk < k1) {
                        uint32_t b0, b1;
                        ldsm_x2_s(b0, b1, wa - (uint32_t)((lane >> 4) * 32));    // x2 uses lanes 0-15 only: drop the k-step kk+1 offset
                        uint32_t a0 = 0, a2 = 0;
                        if (live) { a0 = lds32_s(xa); a2 = lds32_s(xa + 16); }
                        mma_rows8(c, a0, a2, b0, b1);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_s(sm.empty + my_slot * 8);
        }
        sg = 0;
        const float r0 = c[0] + c2[0], r1 = c[1] + c2[1];
        // hand the K-slice partial of this row group to the epilogue warp
        const uint32_t buf = g_idx % MK_NRED;
        mbar_wait_s(sm.red_empty + buf * 8, ((g_idx / MK_NRED) & 1) ^ 1, 5);
        if (live) {
            float* red = sm.red + ((size_t)(buf * MK_CW + warp) * B + g) * 8;
            red[2 * t] = r0;
            red[2 * t + 1] = r1;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_s(sm.red_full + buf * 8);
    }
}

// ---------------------------------------------------------------- attention phase (LL in, LL out)
struct AttnItem { int b, kvh, split, k_begin, k_end, cur, pos; };
struct AttnDims { int nh, nkv, s_max, n_split; const bf16 *cos_tab, *sin_tab; };   // passed BY VALUE to the noinline phase functions: a
                                                                                   // reference to the kernel parameters would force a local copy
constexpr int PART_LD = DA_HD + 4;   // LL words per (row, head, split) partial: o[128], m, l, pad (16-byte aligned rows)

// Same arithmetic as decode_attn_body (decode_common.cuh) -- bf16 M-RoPE semantics, FA2 rounding of P, fp32 statistics -- with the
// q/k/v of this step POLLED from the LL vector the q|k|v epilogues write, and the split partial (m, l, o[128]) published as LL words.
__device__ __forceinline__ void attn_split_ll(const AttnDims p, const AttnItem it, bf16* kc, bf16* vc, const ull* ll_qkv, unsigned int ep_qkv,
                                              ull* ll_part, unsigned int ep_part, uint8_t* smem_raw, int tid, bool first_tile_in_flight,
                                              long long* tl) {
    bf16* sK = reinterpret_cast<bf16*>(smem_raw);           // [64][136]
    long long tl0 = tl ? clock64() : 0;
    bf16* sV = sK + DA_TILE * DA_LD;                        // [64][136]
    float* sQ = reinterpret_cast<float*>(sV + DA_TILE * DA_LD);  // [group][128]
    float* sP = sQ + DA_MAXG * DA_HD;                       // [group][64]
    const int warp = tid >> 5, lane = tid & 31;
    const int group = p.nh / p.nkv;
    const int heads_total = p.nh + 2 * p.nkv;
    const int b = it.b, kvh = it.kvh, cur = it.cur, pos = it.pos, k_begin = it.k_begin, k_end = it.k_end;
    const ull* row = ll_qkv + (size_t)b * (heads_total * DA_HD / 2);
    const float scale_log2 = rsqrtf((float)DA_HD) * 1.4426950408889634f;
    DecAttnArgs da;                                          // only what decode_attn_issue_tile reads
    da.kcache = kc - ((size_t)b * p.nkv + kvh) * p.s_max * DA_HD; da.vcache = vc - ((size_t)b * p.nkv + kvh) * p.s_max * DA_HD;
    da.nkv = p.nkv; da.s_max = p.s_max;

    // ---- roped queries of this kv group -> sQ (fp32 values of the bf16-rounded result); one thread = one (head, column pair)
    for (int i = tid; i < group * 32; i += MK_CTHREADS) {
        const int g = i >> 5, jp = i & 31, j = 2 * jp;
        const ull* qw = row + (size_t)(kvh * group + g) * (DA_HD / 2);
        uint32_t wlo, whi, d0, d1;
        {   // the two words are 256 bytes apart: poll them together
            ull x, y;
            unsigned int spins = 0;
            for (;;) {
                asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(x) : "l"(qw + jp) : "memory");
                asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(y) : "l"(qw + 32 + jp) : "memory");
                if (ll_ok(x, ep_qkv) && ll_ok(y, ep_qkv)) break;
                ll_backoff(); if (++spins > LL_SPIN_LIMIT) ll_timeout(qw, ep_qkv, 20);
            }
            wlo = (uint32_t)x; whi = (uint32_t)y; (void)d0; (void)d1;
        }
        const float2 lo = bf2x_to_f2(wlo), hi = bf2x_to_f2(whi);
        const float c0 = bf2f(p.cos_tab[(size_t)pos * 64 + j]), s0 = bf2f(p.sin_tab[(size_t)pos * 64 + j]);
        const float c1 = bf2f(p.cos_tab[(size_t)pos * 64 + j + 1]), s1 = bf2f(p.sin_tab[(size_t)pos * 64 + j + 1]);
        sQ[g * DA_HD + j] = rbf(rbf(lo.x * c0) + rbf(-hi.x * s0));
        sQ[g * DA_HD + j + 64] = rbf(rbf(hi.x * c0) + rbf(lo.x * s0));
        sQ[g * DA_HD + j + 1] = rbf(rbf(lo.y * c1) + rbf(-hi.y * s1));
        sQ[g * DA_HD + j + 65] = rbf(rbf(hi.y * c1) + rbf(lo.y * s1));
    }
    if (tl) { long long t1 = clock64(); tl[0] += t1 - tl0; tl0 = t1; }      // wait for q
    const bool owns_new = (cur >= k_begin && cur < k_end);
    // the split that owns the key of this step polls its raw k / v together with the queries
    float2 nk_lo = make_float2(0.f, 0.f), nk_hi = nk_lo;
    float nc0 = 0.f, ns0 = 0.f, nc1 = 0.f, ns1 = 0.f;
    uint32_t nv2 = 0;
    if (owns_new) {
        if (tid < 32) {
            const ull* kw = row + (size_t)(p.nh + kvh) * (DA_HD / 2);
            ull x, y;
            unsigned int spins = 0;
            for (;;) {
                asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(x) : "l"(kw + tid) : "memory");
                asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(y) : "l"(kw + 32 + tid) : "memory");
                if (ll_ok(x, ep_qkv) && ll_ok(y, ep_qkv)) break;
                ll_backoff(); if (++spins > LL_SPIN_LIMIT) ll_timeout(kw, ep_qkv, 21);
            }
            nk_lo = bf2x_to_f2((uint32_t)x); nk_hi = bf2x_to_f2((uint32_t)y);
            nc0 = bf2f(p.cos_tab[(size_t)pos * 64 + 2 * tid]); ns0 = bf2f(p.sin_tab[(size_t)pos * 64 + 2 * tid]);
            nc1 = bf2f(p.cos_tab[(size_t)pos * 64 + 2 * tid + 1]); ns1 = bf2f(p.sin_tab[(size_t)pos * 64 + 2 * tid + 1]);
        } else if (tid >= 64 && tid < 128) {
            const ull* vw = row + (size_t)(p.nh + p.nkv + kvh) * (DA_HD / 2) + (tid - 64);
            ull x;
            unsigned int spins = 0;
            for (;;) {
                asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(x) : "l"(vw) : "memory");
                if (ll_ok(x, ep_qkv)) break;
                ll_backoff(); if (++spins > LL_SPIN_LIMIT) ll_timeout(vw, ep_qkv, 22);
            }
            nv2 = (uint32_t)x;
        }
    }

    float m_run = -INFINITY, l_run = 0.f;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int t0 = k_begin; t0 < k_end; t0 += DA_TILE) {
        KSync()();   // previous tile fully consumed (also orders the sQ writes on the first pass)
        if (!(first_tile_in_flight && t0 == k_begin)) decode_attn_issue_tile(da, kvh, b, t0, k_end, cur, smem_raw, tid, MK_CTHREADS);
        cp_async_wait<0>();
        KSync()();
        if (owns_new && cur >= t0 && cur < t0 + DA_TILE) {
            // new key/value of this step: rope k in bf16 semantics, append to the cache, and use it from smem
            const int r = cur - t0;
            if (tid < 32) {
                const int j = 2 * tid;
                const bf16 olo0 = f2bf(rbf(nk_lo.x * nc0) + rbf(-nk_hi.x * ns0)), ohi0 = f2bf(rbf(nk_hi.x * nc0) + rbf(nk_lo.x * ns0));
                const bf16 olo1 = f2bf(rbf(nk_lo.y * nc1) + rbf(-nk_hi.y * ns1)), ohi1 = f2bf(rbf(nk_hi.y * nc1) + rbf(nk_lo.y * ns1));
                sK[r * DA_LD + j] = olo0; sK[r * DA_LD + j + 1] = olo1; sK[r * DA_LD + j + 64] = ohi0; sK[r * DA_LD + j + 65] = ohi1;
                kc[(size_t)cur * DA_HD + j] = olo0; kc[(size_t)cur * DA_HD + j + 1] = olo1;
                kc[(size_t)cur * DA_HD + j + 64] = ohi0; kc[(size_t)cur * DA_HD + j + 65] = ohi1;
            } else if (tid >= 64 && tid < 128) {
                const int j = 2 * (tid - 64);
                *reinterpret_cast<uint32_t*>(sV + r * DA_LD + j) = nv2;
                *reinterpret_cast<uint32_t*>(vc + (size_t)cur * DA_HD + j) = nv2;
            }
            KSync()();
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
                sc[kk] = (t0 + r < k_end) ? d * scale_log2 : -INFINITY;
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
                float pv = pw[r];
                uint2 vv = *reinterpret_cast<const uint2*>(sV + r * DA_LD + lane * 4);
                float2 v0 = bf2x_to_f2(vv.x), v1 = bf2x_to_f2(vv.y);
                acc[0] = fmaf(pv, v0.x, acc[0]); acc[1] = fmaf(pv, v0.y, acc[1]);
                acc[2] = fmaf(pv, v1.x, acc[2]); acc[3] = fmaf(pv, v1.y, acc[3]);
            }
            __syncwarp();
        }
    }
    if (tl) { long long t1 = clock64(); tl[1] += t1 - tl0; tl0 = t1; }      // k/v poll + tiles + softmax
    // ---- publish the partial of this split: 130 LL words per head
    if (warp < group) {
        const int head = kvh * group + warp;
        ull* dst = ll_part + (((size_t)b * p.nh + head) * p.n_split + it.split) * PART_LD;
        ll_store2(dst + lane * 4, ep_part, __float_as_uint(acc[0]), __float_as_uint(acc[1]));
        ll_store2(dst + lane * 4 + 2, ep_part, __float_as_uint(acc[2]), __float_as_uint(acc[3]));
        if (lane == 0) ll_store2(dst + DA_HD, ep_part, __float_as_uint(m_run), __float_as_uint(l_run));
    }
}

// Distributed combine of the split-KV partials of one (row, kv head): the n_split CTAs that produced the partials each reduce
// a slice of the group*128 output values over ALL splits (one warp per 4 values, lanes over splits) as soon as the words arrive.
__device__ __forceinline__ void attn_merge_ll(const AttnDims p, int b, int kvh, int split, int tid, const ull* ll_part, unsigned int ep_part,
                                              ull* ll_attn, unsigned int ep_attn, unsigned int ll_rep, unsigned int ll_nrep) {
    const int warp = tid >> 5, lane = tid & 31;
    const int group = p.nh / p.nkv;
    const int total = group * DA_HD;                                  // values of this kv head's query heads
    int per = (total + p.n_split - 1) / p.n_split;
    per = (per + 3) & ~3;
    const int e_lo = split * per, e_hi = (e_lo + per) < total ? (e_lo + per) : total;
    for (int e0 = e_lo + warp * 4; e0 < e_hi; e0 += (MK_CTHREADS / 32) * 4) {
        const int head = kvh * group + e0 / DA_HD, dim = e0 % DA_HD;
        const ull* base = ll_part + ((size_t)b * p.nh + head) * p.n_split * PART_LD;
        float M = -INFINITY;
        float mj[3], lj[3];
        float4 oj[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {                                  // n_split <= 96 here (launch checks <= 64)
            const int j = lane + 32 * k;
            mj[k] = -INFINITY; lj[k] = 0.f; oj[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j < p.n_split) {
                const ull* pj = base + (size_t)j * PART_LD;
                uint32_t mu, lu;
                ll_wait2(pj + DA_HD, ep_part, mu, lu, 30);
                const uint4 o = ll_wait4(pj + dim, ep_part, 31);
                mj[k] = __uint_as_float(mu); lj[k] = __uint_as_float(lu);
                oj[k] = make_float4(__uint_as_float(o.x), __uint_as_float(o.y), __uint_as_float(o.z), __uint_as_float(o.w));
            }
            M = fmaxf(M, mj[k]);
        }
        M = warp_max(M);
        float L = 0.f, o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float w = (mj[k] == -INFINITY) ? 0.f : exp2f(mj[k] - M);
            L += lj[k] * w;
            o0 += oj[k].x * w; o1 += oj[k].y * w; o2 += oj[k].z * w; o3 += oj[k].w * w;
        }
        L = warp_sum(L); o0 = warp_sum(o0); o1 = warp_sum(o1); o2 = warp_sum(o2); o3 = warp_sum(o3);
        if (lane == 0) {
            float inv = L > 0.f ? 1.0f / L : 0.f;
            ull* dst = ll_attn + (size_t)b * (p.nh * DA_HD / 2) + ((size_t)head * DA_HD + dim) / 2;
            const uint32_t w0 = f2_to_bf2x(o0 * inv, o1 * inv), w1 = f2_to_bf2x(o2 * inv, o3 * inv);
            for (unsigned int r = 0; r < ll_nrep; ++r, dst += ll_rep) ll_store2(dst, ep_attn, w0, w1);
        }
    }
}

template <bool TL>
__global__ void __launch_bounds__(MK_THREADS, 1) decode_mega_kernel(MegaArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const DecodeParams& p = a.p;
    uint8_t* ring = smem;
    uint8_t* xa = ring + (size_t)a.n_slots * a.slot_bytes;    // activation rows; the attention scratch sits at xa + attn_off
    bf16* xs = reinterpret_cast<bf16*>(xa);
    bf16* hraw = reinterpret_cast<bf16*>(xa + a.hraw_off);    // raw residual rows [2][B][D]: copy 0 = input of the layer (residual of o_proj), copy 1 =
                                                              // input of the MLP (residual of down_proj); two copies so the epilogue warp may run behind
    uint64_t* full = reinterpret_cast<uint64_t*>(xa + a.x_bytes);
    uint64_t* empty = full + MK_MAXSLOTS;
    uint64_t* red_full = empty + MK_MAXSLOTS;
    uint64_t* red_empty = red_full + MK_NRED;
    float* red = reinterpret_cast<float*>(red_empty + MK_NRED);          // [MK_NRED][MK_CW][B][8]
    __shared__ float s_pre[8 * MK_MAXGROUPS * 8];
    __shared__ volatile unsigned int s_prog;
    __shared__ long long s_prod_wait;
    __shared__ DecodeLayerW s_w[2];
    __shared__ float s_part[MK_CW];
    __shared__ ArgMax am;
    __shared__ long long s_tok[ST_BMAX], s_arg[ST_BMAX];
    __shared__ int s_unf[ST_BMAX];
    __shared__ const bf16* s_row[ST_BMAX];
    __shared__ volatile unsigned int s_epi_done;
    __shared__ AttnItem s_item;      // rarely used per-launch values live in shared memory: the consumer loops need the registers
    __shared__ int s_plo, s_phi;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = gridDim.x, cta = blockIdx.x;

    if (tid == 0) {
        s_prog = 0; s_prod_wait = 0; s_epi_done = 0;
        for (int i = 0; i < a.n_slots; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], MK_CW); }   // every consumer warp reads every stage
        for (int i = 0; i < MK_NRED; ++i) { mbar_init(&red_full[i], MK_CW); mbar_init(&red_empty[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == MK_CW + 1) {   // producer warp: streams this CTA's weight schedule for the whole step
        producer_loop(a, ring, full, empty, cta, G, lane, &s_prog, &s_prod_wait);
        return;
    }

    if (warp == MK_CW + 2) {   // L2 prefetch warp
        prefetch_loop(a, cta, lane, &s_prog);
        return;
    }

    SmemMap sm;
    sm.ring = smem_u32(ring); sm.full = smem_u32(full); sm.empty = smem_u32(empty);
    sm.red_full = smem_u32(red_full); sm.red_empty = smem_u32(red_empty); sm.xs = smem_u32(xs);
    sm.red = red; sm.pre = s_pre;
    uint32_t it = 0, grp = 0;
    if (warp == MK_CW)
        for (int k = 0; k < 2; ++k) {
            int eb = k * 4 + (lane >> 3);
            am.v[k][lane] = -INFINITY; am.i[k][lane] = 0x7fffffff;
            am.lo[k][lane] = eb < p.B ? p.dev_state[st_vlo(eb)] : 0;
            am.hi[k][lane] = eb < p.B ? p.dev_state[st_vhi(eb)] : 0;
        }
    // cycle accounting (PADT_MEGA_TIMELINE & 1): per phase kind [qkv | attention | o_proj | gate-up | down | logits]
    long long t_work[6] = {0, 0, 0, 0, 0, 0}, t_stage[6] = {0, 0, 0, 0, 0, 0}, t_aux[4] = {0, 0, 0, 0};
    long long t0 = TL ? clock64() : 0, tt = 0;
    long long t_gap[7] = {0, 0, 0, 0, 0, 0, 0}, t_end = t0;   // time outside the buckets, attributed to the phase kind that follows
    int tl_k = 0;
#define TSTART() do { if (TL) { tt = clock64(); t_gap[tl_k] += tt - t_end; } } while (0)
#define TADD(acc) do { if (TL) { t_end = clock64(); acc += t_end - tt; } } while (0)

    if (warp == 0 && lane < (int)(sizeof(DecodeLayerW) / 4)) reinterpret_cast<uint32_t*>(&s_w[0])[lane] = reinterpret_cast<const uint32_t*>(&a.layers[0])[lane];
    // ---- phase 0: sample the token predicted by the previous forward (padt.py:745-757).  EVERY CTA does this redundantly from the
    // arg-max partials of the previous launch (same inputs, same order -> same token), so no CTA waits for another one; the
    // bookkeeping (state block, token logs) is written once, by CTA 0, at the END of the launch -- the state block is read-only
    // while the step runs.
    const int t_step = p.dev_state[0];            // tokens sampled before this launch == index of the token fed by this step
    const int L_prompt = p.dev_state[2];
    if (warp < p.B) {
        const int b = warp;
        float bv = -INFINITY; int bi = 0x7fffffff;
        const int n_part = p.dev_state[3];        // written by whichever kernel produced the arg-max partials
        for (int k = lane; k < n_part; k += 32) {
            float v = p.amax_val[(size_t)k * p.B + b]; int i = p.amax_idx[(size_t)k * p.B + b];
            if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, bv, o); int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) {
            long long arg = bi, tok = arg;
            if (p.forced && t_step < p.max_steps) tok = p.forced[(size_t)t_step * p.B + b];
            int unf = p.dev_state[st_unfin(b)];
            if (p.n_eos > 0) tok = unf ? tok : p.pad_id;
            bool is_eos = false;
            for (int e = 0; e < p.n_eos; ++e) is_eos |= (tok == p.eos_ids[e]);
            if (is_eos) unf = 0;
            s_tok[b] = tok; s_arg[b] = arg; s_unf[b] = unf;
            s_row[b] = (tok < p.vocab) ? p.embed + (size_t)tok * p.D : p.protos + (size_t)(tok - p.vocab) * p.D;   // padt.py:226-229
        }
    }
    CSync()();
    const long long t_sample = TL ? clock64() - t0 : 0;

    const int qkv_n = (p.nh + 2 * p.nkv) * p.hd;
    const int attn_n = p.nh * p.hd;
    const size_t layer_kv = (size_t)p.B * p.nkv * p.s_max * p.hd;
    uint8_t* attn_smem = xa + a.attn_off;
    // this CTA's attention item (n_split * nkv * B <= grid size: at most one item per CTA)
    const bool has_item = cta < p.n_split * p.nkv * p.B && !(a.dbg & 8);
    if (tid == 0) {
        AttnItem item;
        item.split = cta % p.n_split; item.kvh = (cta / p.n_split) % p.nkv; item.b = cta / (p.n_split * p.nkv);
        item.k_begin = item.k_end = item.cur = item.pos = 0;
        if (has_item) {
            // key range of this (row, split): same partition as decode_attn_range; padt.py:270-276: position = cache_position[0] + rope_deltas
            item.cur = L_prompt + t_step;
            const int ks = p.dev_state[st_kvstart(item.b)];
            const int n_keys = item.cur + 1 - ks;
            int chunk = (n_keys + p.n_split - 1) / p.n_split;
            chunk = (chunk + 31) & ~31;
            item.k_begin = ks + item.split * chunk;
            item.k_end = item.k_begin + chunk;
            item.k_end = item.k_end > item.cur + 1 ? item.cur + 1 : item.k_end;
            int pos = item.cur + p.dev_state[st_delta(item.b)];
            item.pos = pos < 0 ? 0 : (pos >= p.max_pos ? p.max_pos - 1 : pos);   // the host sizes the table for L + max_new_tokens (engine.py)
        }
        s_item = item;
        int plo, phi;
        proto_range(p, plo, phi);
        s_plo = plo; s_phi = phi;
    }
    CSync()();
    const AttnItem& item = s_item;
    const unsigned int step_idx = (unsigned int)t_step;                     // decode launches since the prefill

    // ---- ONE phase loop (a single copy of the staging / consume code): 5 phases per layer, then head rows, then the prototype
    //      rows of the request.  Phase ph of this launch tags its LL words with epoch step_idx * n_phase + ph + 1.
    const int n_phase = p.n_layers * 5 + 2;
    bool kv_in_flight = false;
    uint32_t next_w = 0;
    static_assert(sizeof(DecodeLayerW) % 4 == 0 && sizeof(DecodeLayerW) / 4 <= 32, "layer descriptor is copied by one warp");
    for (int ph = 0; ph < n_phase; ++ph) {
        const int li = ph / 5, k = ph < p.n_layers * 5 ? ph % 5 : 5 + (ph - p.n_layers * 5);   // k: 0..4 layer phases, 5 head, 6 prototypes
        const unsigned int epoch = step_idx * (unsigned int)n_phase + (unsigned int)ph + 1u;
        if (TL) tl_k = k;
        ull* llp = a.ll + (size_t)(li & 1) * a.ll_stride;                  // parity copy of this layer (replica 0: what writers address)
        const size_t my_rep = (size_t)((unsigned int)cta % a.ll_nrep) * a.ll_rep;   // the replica this CTA reads
#ifdef PADT_MK_WATCHDOG
        if (tid == 0) { WD_MARK(g_wd_phase, ph); WD_MARK(g_wd_sub, 0); }
#endif
        // next layer's descriptor: the global load is issued at the start of the layer and parked in a register, the shared-memory store
        // happens three phases later -- storing right after the load made warp 0 (and with it every staging barrier) wait ~1.4 us per layer
        if (warp == 0 && lane < (int)(sizeof(DecodeLayerW) / 4) && li + 1 < p.n_layers) {
            if (k == 0) next_w = reinterpret_cast<const uint32_t*>(&a.layers[li + 1])[lane];
            if (k == 3) reinterpret_cast<uint32_t*>(&s_w[(li + 1) & 1])[lane] = next_w;
        }
        const DecodeLayerW& w = s_w[li & 1];       // fields are read at their point of use (keeps registers free)
        if (k == 1) {
            // ---- M-RoPE + KV append + split-KV attention (consumer warps); the n_split CTAs of a (row, kv head) then combine slices of the result
            TSTART();
            if (has_item && warp < MK_CW) {
                bf16* kc = p.kcache + (size_t)li * layer_kv + ((size_t)item.b * p.nkv + item.kvh) * p.s_max * DA_HD;
                bf16* vc = p.vcache + (size_t)li * layer_kv + ((size_t)item.b * p.nkv + item.kvh) * p.s_max * DA_HD;
                AttnDims ad;
                ad.nh = p.nh; ad.nkv = p.nkv; ad.s_max = p.s_max; ad.n_split = p.n_split; ad.cos_tab = p.cos_tab; ad.sin_tab = p.sin_tab;
                attn_split_ll(ad, item, kc, vc, llp + my_rep + a.ll_qkv, epoch - 1, llp + a.ll_part, epoch, attn_smem, tid, kv_in_flight,
                              TL ? t_aux : nullptr);
                long long tm0 = TL ? clock64() : 0;
                attn_merge_ll(ad, item.b, item.kvh, item.split, tid, llp + a.ll_part, epoch, llp + a.ll_attn, epoch, a.ll_rep, a.ll_nrep);
                if (TL) t_aux[2] += clock64() - tm0;
            }
            TADD(t_work[1]);
            continue;
        }
        // ---- projection phases: the switch only picks parameters; ONE copy of the staging code follows
        Seg s;
        const bf16* bias = nullptr;
        ull* ll_out = nullptr;
        int ep, n_out = 0;
        unsigned int wait_layers = 0;
        const int tk = k < 5 ? k : 5;
        const ull* st_ll = nullptr;         // staging source (LL words) ...
        const bf16* st_ln = nullptr;        // ... RMSNorm weight (null: plain copy)
        int st_K = p.D;
        bool st_raw = false, st_on = true;
        TSTART();
        switch (k) {
            case 0:
                s = layer_seg(a, w, SEG_QKV, cta, G); ep = EP_BIAS; bias = w.bqkv; ll_out = llp + a.ll_qkv; n_out = qkv_n;
                wait_layers = (unsigned int)li;
                // residual stream: layer 0 reads the embedding row of the sampled token, later layers the rows the previous down_proj published
                st_ll = li == 0 ? nullptr : a.ll + (size_t)((li - 1) & 1) * a.ll_stride + my_rep + a.ll_hA;
                st_ln = w.ln1; st_raw = true;
                break;
            case 2:
                s = layer_seg(a, w, SEG_O, cta, G); ep = EP_RESID; ll_out = llp + a.ll_hB; n_out = p.D;
                st_ll = llp + my_rep + a.ll_attn; st_K = attn_n;
                break;
            case 3:
                s = layer_seg(a, w, SEG_GU, cta, G); ep = EP_SWIGLU; ll_out = llp + a.ll_act; n_out = p.I;
                st_ll = llp + my_rep + a.ll_hB; st_ln = w.ln2; st_raw = true;
                break;
            case 4:
                s = layer_seg(a, w, SEG_DOWN, cta, G); ep = EP_RESID; ll_out = llp + a.ll_hA; n_out = p.D;
                st_ll = llp + my_rep + a.ll_act; st_K = p.I;
                break;
            case 5:
                s = make_seg(p.head, p.head_img, p.nseg_d, 0, p.vocab, p.D, cta, G, a.slot_bytes);
                ep = EP_LOGITS;
                st_ll = a.ll + (size_t)((p.n_layers - 1) & 1) * a.ll_stride + my_rep + a.ll_hA; st_ln = p.final_norm;
                break;
            default:
                if (s_phi <= s_plo) continue;
                s = make_seg(p.protos + (size_t)s_plo * p.D, nullptr, 0, p.vocab + s_plo, s_phi - s_plo, p.D, cta, G, a.slot_bytes);
                ep = EP_LOGITS; st_on = false;      // the normalised rows are still staged
                break;
        }
        if (st_on) {
            if (st_ln) {
                if (st_raw && warp < MK_CW) {
                    // the raw residual rows about to be overwritten were last read by the epilogue warp two phases ago; it runs behind the
                    // consumers by at most MK_NRED row groups, so this is (almost) never a wait -- it only makes the reuse safe for CTAs that
                    // own no rows in the phases in between (miniature models)
                    unsigned int spins = 0;
                    while (s_epi_done < grp)
                        if (++spins > LL_SPIN_LIMIT) ll_timeout((const void*)&s_epi_done, grp, 15);
                }
                bf16 *copy0 = nullptr, *copy1 = nullptr;
                if (k == 5 && cta == 0) {
                    copy0 = p.last_hidden;
                    const int t = t_step + 1;     // tokens sampled so far, this launch's sample included
                    if (p.hidden_steps && t < p.max_steps) copy1 = p.hidden_steps + (size_t)t * p.B * p.D;
                }
                stage_norm(s_row, st_ll, epoch - 1, st_ln, p.eps, p.D, p.B, xs, st_raw ? hraw + (k == 3 ? (size_t)p.B * p.D : 0) : nullptr, tid, lane, copy0,
                           copy1, s_part);
            } else {
                stage_copy(st_ll, epoch - 1, st_K, p.B, xs, tid);
            }
            if (k == 0) {
                kv_in_flight = false;
                if (has_item && item.k_end > item.k_begin && (a.flags & MF_KV_SMEM)) {
                    // the cache rows of earlier steps do not depend on this step: pull the item's first K/V tile into shared memory
                    // now, so that it is resident when the q/k/v words arrive
                    DecAttnArgs da;
                    da.kcache = p.kcache + (size_t)li * layer_kv; da.vcache = p.vcache + (size_t)li * layer_kv; da.nkv = p.nkv; da.s_max = p.s_max;
                        if (warp < MK_CW) decode_attn_issue_tile(da, item.kvh, item.b, item.k_begin, item.k_end, item.cur, attn_smem, tid, MK_CTHREADS);
                    kv_in_flight = true;
                }
            }
        }
        if (k <= 5 && warp < MK_CW) KSync()();      // the epilogue warp never waits here: it sees the staged rows through the reduction mbarriers
        if (k == 0 && tid == 0)
            // "this CTA has finished reading layer li - 1" (WAR guard of the LL parity copies, see consume_seg); release orders the
            // staging loads of thread 0's CTA (bar.sync above) before the increment
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(a.gbar) : "memory");
        TADD(t_stage[tk]);
        if (TL) { tt = clock64(); t_end = tt; }
        if (tid == 0) WD_MARK(g_wd_sub, 1);
        {
            ConsumeArgs c;
            c.s = s; c.sm = sm; c.am = &am; c.bias = bias; c.ll_out = ll_out; c.hraw = hraw + (k == 4 ? (size_t)p.B * p.D : 0); c.xpart = p.xpart; c.gbar = a.gbar;
            c.it = it; c.grp = grp; c.ep = ep; c.tid = tid; c.n_out = n_out; c.B = p.B; c.vocab = p.vocab; c.n_slots = a.n_slots;
            c.slot_bytes = a.slot_bytes; c.flags = a.flags; c.dbg = a.dbg; c.epoch = epoch; c.wait_layers = wait_layers;
            c.ll_rep = a.ll_rep; c.ll_nrep = a.ll_nrep; c.epi_done = &s_epi_done;
            consume_seg(c);
            const int n_st = s.st_hi - s.st_lo;                       // every thread advances the CTA-wide stage / row-group counters identically
            it += (uint32_t)n_st;
            grp += (uint32_t)(n_st > 0 ? (s.st_hi - 1) / s.nseg - s.st_lo / s.nseg + 1 : 0);
        }
        if (tid == 0) WD_MARK(g_wd_sub, 2);
        if (warp < MK_CW) KSync()();     // the activation rows are rewritten by the next staging (the epilogue warp does not read them)
        TADD(t_work[tk]);
    }
    // ---- masked arg-max partial of this CTA
    CSync()();             // the epilogue warp has finished the last row group
    if (tid < p.B) {       // entry [pass][lane]: batch row pass*4 + lane/8, weight-row lane%8
        float bv = -INFINITY; int bi = 0x7fffffff;
        for (int k = 0; k < 8; ++k) {
            float v = am.v[tid >> 2][(tid & 3) * 8 + k]; int i = am.i[tid >> 2][(tid & 3) * 8 + k];
            if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
        }
        p.amax_val[(size_t)cta * p.B + tid] = bv;
        p.amax_idx[(size_t)cta * p.B + tid] = bi;
    }
    if (cta == 0) {
        // ---- bookkeeping of the sample phase, written once the state block is no longer read by anyone in this launch: every CTA
        // has passed the q/k/v staging of the last layer (it read the state block before its first staging)
        if (tid == 0) {
            unsigned int spins = 0;
            while (ld_acquire_u32(a.gbar) < (unsigned int)p.n_layers * (unsigned int)G)
                ll_backoff(); if (++spins > LL_SPIN_LIMIT) ll_timeout(a.gbar, 0, 14);
            int n = 0;
            for (int b = 0; b < p.B; ++b) {
                n += s_unf[b];
                p.dev_state[st_unfin(b)] = s_unf[b];
                if (t_step < p.max_steps) {
                    p.tokens_out[(size_t)t_step * p.B + b] = s_tok[b];
                    if (p.argmax_out) p.argmax_out[(size_t)t_step * p.B + b] = s_arg[b];
                }
                p.cur_token[b] = s_tok[b];
            }
            p.dev_state[1] = n;
            p.dev_state[0] = t_step + 1;
            p.dev_state[3] = G;
        }
    }
    if (TL && (a.dbg & 32) && tid == 0) {
        unsigned int smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        printf("[mega skew] cta %d smid %u work qkv %lld o %lld gu %lld down %lld head %lld | poll qkv %lld o %lld gu %lld down %lld\n", cta, smid, t_work[0], t_work[2],
               t_work[3], t_work[4], t_work[5], t_stage[0], t_stage[2], t_stage[3], t_stage[4]);
    }
    if (TL && (a.dbg & 1) && tid == 0 && (cta == 0 || cta == G / 2 || cta == G - 1)) {
        printf("[mega cta %d] cycles: total %lld | sample %lld\n"
               "[mega cta %d] per phase (qkv attn o gu down logits): work %lld %lld %lld %lld %lld %lld | stage+poll %lld %lld %lld %lld %lld %lld | attn: wait q %lld tiles %lld merge %lld\n",
               cta, (long long)(clock64() - t0), t_sample, cta, t_work[0], t_work[1], t_work[2], t_work[3], t_work[4], t_work[5], t_stage[0],
               t_stage[1], t_stage[2], t_stage[3], t_stage[4], t_stage[5], t_aux[0], t_aux[1], t_aux[2]);
        printf("[mega cta %d] outside the buckets, by following phase kind: %lld %lld %lld %lld %lld %lld %lld\n", cta, t_gap[0], t_gap[1], t_gap[2], t_gap[3],
               t_gap[4], t_gap[5], t_gap[6]);
    }
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

// Largest stage: 8 rows x 1024 columns (+ pad) = 16.5 KB, the stage of every K = 2048 matrix (q/k/v, o_proj, gate/up, LM head at 3B: 71 % of
// the bytes).  Ring slots have ONE size, so a larger cap (24 KB: 22 KB down_proj stages) left 25 % of every slot empty during those phases;
// the stream is latency-bound by the bytes the ring keeps in flight (HBM latency under load x 44 GB/s per SM ~ 90-130 KB).
static int stage_cap() { return getenv("PADT_MEGA_STAGE_CAP") ? atoi(getenv("PADT_MEGA_STAGE_CAP")) : 24 * 1024; }
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

static int decode_mega_attn_off(const DecodeParams& p) { return (p.B * (p.D + 8) * 2 + 127) / 128 * 128; }
// raw residual rows [B][D] sit behind the widest activation rows and the attention scratch (both are live while they are)
static int decode_mega_hraw_off(const DecodeParams& p) {
    int kmax = p.I > p.D ? p.I : p.D;
    kmax = kmax > p.nh * p.hd ? kmax : p.nh * p.hd;
    int xb = p.B * (kmax + 8) * 2;
    const int attn_need = decode_mega_attn_off(p) + MK_ATTN_SMEM;
    xb = xb > attn_need ? xb : attn_need;
    return (xb + 127) / 128 * 128;
}

// Plans shared memory for the persistent kernel; returns 0 and fills the outputs, or -1 when the model does not fit.
int decode_mega_plan(const DecodeParams& p, int* slot_bytes, int* n_slots, int* x_bytes, size_t* smem_bytes) {
    if (p.B > 8) return -1;
    if (p.D > 16 * MK_CTHREADS || p.D % 16 || p.I % 16 || (p.nh * p.hd) % 16) return -1;   // one 8-value chunk of a residual row per staging thread
    // the attention scratch sits behind the activation rows of the q/k/v phase: its first K/V tile is filled while they are in use;
    // the raw residual rows (epilogue residual source) follow the widest activation rows
    int xb = decode_mega_hraw_off(p) + 2 * p.B * p.D * 2;
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
    if (ns < 4) return -1;
    *slot_bytes = sb; *n_slots = ns; *x_bytes = xb;
    *smem_bytes = (size_t)ns * sb + fixed;
    return 0;
}

// Host mirror of make_seg/layer_seg: flattens the image-backed stages of every CTA (same order as the consumers walk them).
// Stage range [lo, hi) of CTA `c` for a matrix of n_rows x K (stage = row group * nseg + K segment); host mirror of make_seg.
void decode_mega_stage_range(int n_rows, int K, int G, bool balanced, int c, int* nseg_out, unsigned int* lo, unsigned int* hi) {
    const int nseg = decode_mega_nseg(K);
    const unsigned int groups = (unsigned int)(n_rows + MK_GROUP - 1) / MK_GROUP;
    if (balanced && groups >= (unsigned int)G) {
        const unsigned int S = groups * (unsigned int)nseg;
        *lo = (S * (unsigned int)c) / (unsigned int)G; *hi = (S * (unsigned int)(c + 1)) / (unsigned int)G;
    } else {
        *lo = ((groups * (unsigned int)c) / (unsigned int)G) * nseg; *hi = ((groups * (unsigned int)(c + 1)) / (unsigned int)G) * nseg;
    }
    *nseg_out = nseg;
}

int decode_mega_build_table(const DecodeParams& p, const DecodeLayerW* layers_host, int G, bool split, std::vector<uint64_t>& tab, std::vector<int>& off) {
    tab.clear();
    off.assign(G + 1, 0);
    for (int c = 0; c < G; ++c) {
        off[c] = (int)(tab.size() / 2);
        auto emit = [&](const bf16* img, int n_rows, int K, bool balanced) {
            int nseg;
            unsigned int lo, hi;
            decode_mega_stage_range(n_rows, K, G, balanced, c, &nseg, &lo, &hi);
            const uint32_t bytes = MK_GROUP * ((uint32_t)(K / nseg) * 2 + 16);
            for (unsigned int st = lo; st < hi; ++st) {
                const uint8_t* src = reinterpret_cast<const uint8_t*>(img) + (size_t)st * bytes;
                tab.push_back((uint64_t)(uintptr_t)src);
                tab.push_back((uint64_t)bytes);
            }
        };
        for (int li = 0; li < p.n_layers; ++li) {
            const DecodeLayerW& w = layers_host[li];
            emit(w.iqkv, (p.nh + 2 * p.nkv) * p.hd, p.D, false);
            emit(w.io, p.D, p.nh * p.hd, false);
            emit(w.igu, 2 * p.I, p.D, false);
            emit(w.idown, p.D, p.I, split);        // see layer_seg
        }
        emit(p.head_img, p.vocab, p.D, false);
    }
    off[G] = (int)(tab.size() / 2);
    return 0;
}

// LL exchange buffers: word offsets inside one parity copy (see MegaArgs).  off[0..4]: hA, hB, qkv, attn, act inside a replica;
// off[5]: the split-KV partials (not replicated: each word has <= 8 readers), behind `nrep` replicas; off[6]: words per replica.
constexpr int LL_MAXREP = 16;
size_t decode_mega_ll_words(const DecodeParams& p, int nrep, unsigned int* off7) {
    const size_t qkv_n = (size_t)(p.nh + 2 * p.nkv) * p.hd, attn_n = (size_t)p.nh * p.hd;
    size_t o = 0;
    auto take = [&](size_t words) { size_t r = o; o = (o + words + 15) / 16 * 16; return (unsigned int)r; };   // 128-byte aligned vectors
    off7[0] = take((size_t)p.B * p.D / 2);                        // hA: output of down_proj (+ residual)
    off7[1] = take((size_t)p.B * p.D / 2);                        // hB: output of o_proj (+ residual)
    off7[2] = take((size_t)p.B * qkv_n / 2);
    off7[3] = take((size_t)p.B * attn_n / 2);
    off7[4] = take((size_t)p.B * p.I / 2);
    off7[6] = (unsigned int)o;
    o *= (size_t)nrep;
    off7[5] = take((size_t)p.B * p.nh * p.n_split * PART_LD);
    return o;                                                     // words per parity copy; the buffer holds two
}
size_t decode_mega_ll_alloc_words(const DecodeParams& p) {
    unsigned int off7[7];
    return 2 * decode_mega_ll_words(p, LL_MAXREP, off7);
}

int decode_mega_flags() { return getenv("PADT_MEGA_FLAGS") ? atoi(getenv("PADT_MEGA_FLAGS")) : MK_DEFAULT_FLAGS; }
bool decode_mega_split() { return (decode_mega_flags() & MF_SPLIT) != 0; }

// Every precondition of the persistent kernel in one predicate: run_step falls back to the per-layer kernels (decode.cu) when it says no.
bool decode_mega_can_run(const DecodeParams& p, int n_sms) {
    int sb, ns, xb;
    size_t smem;
    if (decode_mega_plan(p, &sb, &ns, &xb, &smem)) return false;
    if (p.hd != DA_HD || p.nh % p.nkv || p.nh / p.nkv > DA_MAXG) return false;
    if (p.n_split > 64 || p.n_split * p.nkv * p.B > n_sms) return false;                       // one attention item per CTA
    const int qkv_n = (p.nh + 2 * p.nkv) * p.hd;
    if (qkv_n % MK_GROUP || p.D % MK_GROUP || (2 * p.I) % MK_GROUP) return false;            // whole row groups (LL words carry row pairs)
    if ((qkv_n / MK_GROUP + n_sms - 1) / n_sms + 2 > MK_MAXGROUPS) return false;              // bias preload buffer
    if (((uintptr_t)p.head & 15) || (p.protos && ((uintptr_t)p.protos & 15))) return false;   // TMA bulk copies need 16-byte aligned rows
    int dev = 0, coop = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) != cudaSuccess || !coop) return false;
    return true;
}

int launch_decode_mega(const DecodeParams& p, const DecodeLayerW* layers_dev, const void* stage_tab, const int* stage_off, unsigned int* gbar,
                       int n_sms, cudaStream_t st) {
    PADT_REQUIRE(decode_mega_can_run(p, n_sms), "persistent decode kernel: configuration not supported (callers check decode_mega_can_run)");
    MegaArgs a;
    a.p = p; a.layers = layers_dev; a.gbar = gbar;
    a.stage_tab = reinterpret_cast<const StageDesc*>(stage_tab); a.stage_off = stage_off;
    a.dbg = getenv("PADT_MEGA_TIMELINE") ? atoi(getenv("PADT_MEGA_TIMELINE")) : 0;
    a.pf_depth = getenv("PADT_MEGA_PF") ? atoi(getenv("PADT_MEGA_PF")) : MK_DEFAULT_PF;
    a.flags = decode_mega_flags();
    a.pf_pace = getenv("PADT_MEGA_PF_PACE") ? atoi(getenv("PADT_MEGA_PF_PACE")) : MK_DEFAULT_PF_PACE;
    a.attn_off = decode_mega_attn_off(p);
    a.hraw_off = decode_mega_hraw_off(p);
    {
        unsigned int off7[7];
        int nrep = getenv("PADT_MEGA_LL_REP") ? atoi(getenv("PADT_MEGA_LL_REP")) : LL_MAXREP;
        nrep = nrep < 1 ? 1 : (nrep > LL_MAXREP ? LL_MAXREP : nrep);
        const size_t words = decode_mega_ll_words(p, nrep, off7);
        PADT_REQUIRE(p.ll != nullptr && p.ll_words >= 2 * words && words < (1u << 31), "LL exchange buffer");
        a.ll = p.ll; a.ll_stride = (unsigned int)words;
        a.ll_hA = off7[0]; a.ll_hB = off7[1]; a.ll_qkv = off7[2]; a.ll_attn = off7[3]; a.ll_act = off7[4]; a.ll_part = off7[5];
        a.ll_rep = off7[6]; a.ll_nrep = (unsigned int)nrep;
    }
    size_t smem;
    if (decode_mega_plan(p, &a.slot_bytes, &a.n_slots, &a.x_bytes, &smem)) {
        padt_set_error("persistent decode kernel: configuration does not fit in shared memory");
        return -1;
    }
    {   // the opt-in shared-memory size is a per-DEVICE function attribute: track it per device (several engines in one process)
        static size_t configured[64] = {0};
        int dev = 0;
        PADT_CHECK_CUDA(cudaGetDevice(&dev));
        if (dev < 0 || dev >= 64 || smem > configured[dev]) {
            PADT_CHECK_CUDA(cudaFuncSetAttribute(decode_mega_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            PADT_CHECK_CUDA(cudaFuncSetAttribute(decode_mega_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            if (dev >= 0 && dev < 64) configured[dev] = smem;
        }
    }
    {
        static int sleep_set = -1;
        const int want = getenv("PADT_MEGA_LL_SLEEP") ? atoi(getenv("PADT_MEGA_LL_SLEEP")) : 0;
        if (want != sleep_set) {
            cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
            cudaStreamIsCapturing(st, &cs);
            if (cs == cudaStreamCaptureStatusNone) { PADT_CHECK_CUDA(cudaMemcpyToSymbol(g_ll_sleep, &want, sizeof(int))); sleep_set = want; }
        }
    }
    PADT_CHECK_CUDA(cudaMemsetAsync(a.gbar, 0, sizeof(unsigned int), st));
    void* args[] = {(void*)&a};
    const void* fn = a.dbg ? (const void*)decode_mega_kernel<true> : (const void*)decode_mega_kernel<false>;   // <true>: cycle accounting
    PADT_CHECK_CUDA(cudaLaunchCooperativeKernel(fn, dim3(n_sms), dim3(MK_THREADS), args, smem, st));
    ++g_padt_launches;
    return 0;
}
