// Persistent decode-step kernel ("weight-streaming megakernel"): ONE launch per generated token.
//
// Organised for the HBM roofline of the step (6.27 GB per token at 3B, B=1); DESIGN.md §4.1 has the measurements.
//   * grid = one CTA per SM, resident for the whole step: 16 tensor-core consumer warps, 1 epilogue warp, 1 TMA producer warp,
//     1 L2 prefetch warp.  Phases (sample | per layer: qkv, attention, o_proj, gate/up, down | head rows | prototype rows) run in
//     ONE phase loop and are separated by a grid barrier instead of kernel boundaries;
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

constexpr int MK_CW = 16;                     // consumer (tensor-core) warps
constexpr int MK_CTHREADS = MK_CW * 32;
constexpr int MK_STHREADS = MK_CTHREADS + 32; // + 1 epilogue warp: the threads that take part in phase-level barriers
constexpr int MK_THREADS = MK_STHREADS + 64;  // + 1 producer warp + 1 L2 prefetch warp
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
};
enum { MF_EVICT_FIRST = 1, MF_NO_AREG = 2, MF_KV_SMEM = 4, MF_PF_LINES = 8, MF_PF_EVICT_LAST = 16, MF_STAGE_MIN = 64, MF_SPLIT = 128 };

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
// Grid barrier over the consumer threads of all CTAs.  bar.sync orders this CTA's writes before thread 0's gpu-scope
// release-increment (cumulativity); the acquire poll + bar.sync publishes the other CTAs' writes to every thread.
__device__ __forceinline__ void grid_sync(unsigned int* bar, unsigned int& target, int G, int tid, int dbg) {
    CSync()();
    if (dbg & 4) return;
    target += (unsigned int)G;
    if (tid == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
#ifdef PADT_MK_WATCHDOG
        WD_DECL(6);
        while (ld_acquire_u32(bar) < target) WD_CHECK();
#else
        while (ld_acquire_u32(bar) < target) {}
#endif
    }
    CSync()();
}

// activation rows -> shared memory, row stride K + 8 elements (the 16-byte pad keeps the A-fragment loads conflict free).
// ln_w != null: RMSNorm prologue, bit-identical to rmsnorm_kernel / the decode.cu prologue (rms_rstd_warp + rms_apply8).
__device__ __forceinline__ void stage_x(const bf16* src, const bf16* ln_w, float eps, int K, int B, bf16* xs, int tid, int lane, bf16* copy0,
                                        bf16* copy1, float* s_part, bool min_traffic) {
    if (tid >= MK_CTHREADS) return;
    const int x_ld = K + 8;
    for (int b = 0; b < B; ++b) {
        const bf16* xr = src + (size_t)b * K;
        if (ln_w && min_traffic && K <= 8 * MK_CTHREADS) {
            // every element of the row is loaded by exactly ONE thread of the CTA (the warp-redundant statistics pass below pulls
            // 16 x the row through L2 -> L1 in all 148 CTAs at once, right after a barrier has invalidated L1): per-warp partial
            // sums of squares -> shared memory -> fixed-order total, then the thread normalises the chunk it already holds.
            const int nch = K >> 3, warp = tid >> 5;
            const bool mine = tid < nch;
            uint4 xv = make_uint4(0, 0, 0, 0), wv = xv;
            if (mine) { xv = *reinterpret_cast<const uint4*>(xr + tid * 8); wv = *reinterpret_cast<const uint4*>(ln_w + tid * 8); }
            float2 e0 = bf2x_to_f2(xv.x), e1 = bf2x_to_f2(xv.y), e2 = bf2x_to_f2(xv.z), e3 = bf2x_to_f2(xv.w);
            float ss = e0.x * e0.x + e0.y * e0.y + e1.x * e1.x + e1.y * e1.y + e2.x * e2.x + e2.y * e2.y + e3.x * e3.x + e3.y * e3.y;
            ss = warp_sum(ss);
            if (lane == 0) s_part[warp] = ss;
            asm volatile("bar.sync 2, %0;" ::"n"(MK_CTHREADS) : "memory");
            float tot = 0.f;
#pragma unroll
            for (int w = 0; w < MK_CW; ++w) tot += s_part[w];
            const float rstd = 1.0f / sqrtf(tot / (float)K + eps);
            if (mine) {
                uint4 o = rms_apply8(xv, wv, rstd);
                *reinterpret_cast<uint4*>(xs + (size_t)b * x_ld + tid * 8) = o;
                if (copy0) *reinterpret_cast<uint4*>(copy0 + (size_t)b * K + tid * 8) = o;
                if (copy1) *reinterpret_cast<uint4*>(copy1 + (size_t)b * K + tid * 8) = o;
            }
            if (b + 1 < B) asm volatile("bar.sync 2, %0;" ::"n"(MK_CTHREADS) : "memory");   // s_part is reused by the next row
        } else if (ln_w) {
            // the norm weights do not depend on the statistics: fetch this thread's first chunk before the reduction, so the
            // apply pass does not start with a second dependent L2 round trip
            uint4 w0 = make_uint4(0, 0, 0, 0);
            if (tid * 8 < K) w0 = *reinterpret_cast<const uint4*>(ln_w + tid * 8);
            float rstd = rms_rstd_warp(xr, K, eps, lane);
            for (int c = tid * 8; c < K; c += MK_CTHREADS * 8) {
                uint4 o = rms_apply8(*reinterpret_cast<const uint4*>(xr + c), c == tid * 8 ? w0 : *reinterpret_cast<const uint4*>(ln_w + c), rstd);
                *reinterpret_cast<uint4*>(xs + (size_t)b * x_ld + c) = o;
                if (copy0) *reinterpret_cast<uint4*>(copy0 + (size_t)b * K + c) = o;
                if (copy1) *reinterpret_cast<uint4*>(copy1 + (size_t)b * K + c) = o;
            }
        } else {
            for (int c = tid * 8; c < K; c += MK_CTHREADS * 8)
                *reinterpret_cast<uint4*>(xs + (size_t)b * x_ld + c) = *reinterpret_cast<const uint4*>(xr + c);
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
__device__ __forceinline__ void consume_seg(const MegaArgs& a, const Seg& s, int ep, const SmemMap& sm, uint32_t& it, uint32_t& grp, int tid,
                                            const bf16* bias, bf16* y, int n_out, ArgMax& am, unsigned int epoch) {
    const DecodeParams& p = a.p;
    const int warp = tid >> 5, lane = tid & 31;
    const int B = p.B;
    const int n_slots = a.n_slots;
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
        if (ep == EP_BIAS || ep == EP_RESID) {   // preload so that the per-group epilogue never waits on global memory
            for (int pass = 0; pass * 4 < B; ++pass) {
                int eb = pass * 4 + (lane >> 3);
                if (eb < B)
                    for (int gi = tail_first ? 1 : 0; gi < n_grp && gi < MK_MAXGROUPS; ++gi) {
                        int r = (g_first + gi) * 8 + en;
                        if (r < s.n_rows)
                            sm.pre[eb * (MK_MAXGROUPS * 8) + gi * 8 + en] =
                                ep == EP_BIAS ? (bias ? bf2f(bias[s.row_begin + r]) : 0.f) : bf2f(y[(size_t)eb * n_out + s.row_begin + r]);
                    }
            }
            __syncwarp();
        }
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
                        const unsigned long long* xp = p.xpart + ((size_t)(blockIdx.x + 1) * 8 + eb) * 8;
                        unsigned long long w0, w1;
                        do {
                            asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w0) : "l"(xp + en) : "memory");
                            asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w1) : "l"(xp + (en | 1)) : "memory");
                        } while ((unsigned int)(w0 >> 32) != epoch || (unsigned int)(w1 >> 32) != epoch);
                        nb0[pass] = __uint_as_float((unsigned int)w0);
                        nb1[pass] = __uint_as_float((unsigned int)w1);
                    }
                }
            }
            mbar_wait_s(sm.red_full + buf * 8, (g_idx / MK_NRED) & 1, 3);
            const float* red = sm.red + (size_t)buf * MK_CW * B * 8;
            const int r = (g_first + gi) * 8 + en;
            if (!(a.dbg & 2)) {
                const int n = s.row_begin + r;
#pragma unroll
                for (int pass = 0; pass < 2; ++pass) {              // static indices keep the arg-max state in registers
                    const int eb = pass * 4 + (lane >> 3);
                    if (eb >= B) break;
                    float sum = 0.f;
#pragma unroll
                    for (int w = 0; w < MK_CW; ++w) sum += red[(w * B + eb) * 8 + en];
                    if (is_tail) {                                  // publish: the neighbour finishes this group
                        const unsigned long long w = ((unsigned long long)epoch << 32) | (unsigned long long)__float_as_uint(sum);
                        asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p.xpart + ((size_t)blockIdx.x * 8 + eb) * 8 + en), "l"(w) : "memory");
                        continue;
                    }
                    if (r >= s.n_rows) continue;
                    sum += nb0[pass];
                    if (ep == EP_BIAS) {
                        y[(size_t)eb * n_out + n] = f2bf(sum + sm.pre[eb * (MK_MAXGROUPS * 8) + gi * 8 + en]);
                    } else if (ep == EP_RESID) {
                        y[(size_t)eb * n_out + n] = f2bf(sm.pre[eb * (MK_MAXGROUPS * 8) + gi * 8 + en] + rbf(sum));
                    } else if (ep == EP_SWIGLU) {
                        if ((en & 1) == 0) {   // even row = gate_j, odd row = up_j of the interleaved weight
                            float up = nb1[pass];
#pragma unroll
                            for (int w = 0; w < MK_CW; ++w) up += red[(w * B + eb) * 8 + en + 1];
                            float gv = rbf(sum), uv = rbf(up);
                            y[(size_t)eb * n_out + (n >> 1)] = f2bf(rbf(silu_f(gv)) * uv);
                        }
                    } else {   // EP_LOGITS: bf16-rounded logit, -inf outside the row's own VRT range, first max wins
                        float l0 = rbf(sum);
                        bool ok = n < p.vocab || (n - p.vocab >= am.lo[pass][lane] && n - p.vocab < am.hi[pass][lane]);
                        const float bv = am.v[pass][lane];
                        if (ok && (l0 > bv || (l0 == bv && n < am.i[pass][lane]))) { am.v[pass][lane] = l0; am.i[pass][lane] = n; }
                        if (p.logits_out) p.logits_out[(size_t)eb * p.logits_ld + n] = f2bf(ok ? l0 : -INFINITY);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive_s(sm.red_empty + buf * 8);
                WD_MARK(g_wd_epi, (int)g_idx);
            }
        }
        return;
    }
    // ================= consumer warps
    const int g = lane >> 2, t = lane & 3;
    const uint32_t row_stride = (uint32_t)s.Kseg * 2 + 16;
    const int set = warp >> 3, wi = warp & 7;
    const int ksteps = s.Kseg >> 4;
    const int k0 = (ksteps * wi) / 8, k1 = (ksteps * (wi + 1)) / 8;
    const bool live = g < B;
    // per-lane byte offset of the ldmatrix.x4 rows inside a stage: matrices 0/1 = k-step kk (k 0-7 / 8-15), 2/3 = k-step kk+1
    const uint32_t w_off = (uint32_t)(lane & 7) * row_stride + (uint32_t)(((lane >> 4) * 16 + ((lane >> 3) & 1) * 8) * 2) + (uint32_t)k0 * 32u;
    const uint32_t x_row = sm.xs + (uint32_t)((g * (s.K + 8) + 2 * t + k0 * 16) * 2);
    const bool reg_path = s.nseg <= 2 && ksteps == 64 && !(a.flags & MF_NO_AREG);
    uint32_t areg[16];
    if (reg_path) {
        // this warp's activation fragments: K segment `set` (segment 0 when there is only one), k-steps [8 wi, 8 wi + 8)
        const uint32_t xb = x_row + (uint32_t)((s.nseg == 2 ? set : 0) * s.Kseg * 2);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            areg[2 * i] = live ? lds32_s(xb + i * 32) : 0u;
            areg[2 * i + 1] = live ? lds32_s(xb + i * 32 + 16) : 0u;
        }
    }
    const bool skip_math = (a.dbg & (2 | 64)) != 0;
    int sg = s.st_lo % s.nseg;                                              // the first group may start in the middle (balanced partition)
    for (int idx = s.st_lo; idx < s.st_hi; ++g_idx) {
        float c[4] = {0.f, 0.f, 0.f, 0.f}, c2[4] = {0.f, 0.f, 0.f, 0.f};
        for (; sg < s.nseg && idx < s.st_hi; ++sg, ++idx) {
            const uint32_t my_slot = slot, my_parity = parity;
            if (++slot == (uint32_t)n_slots) { slot = 0; parity ^= 1u; }
            if (s.nseg > 1 ? ((sg & 1) != set) : (set != 0)) continue;          // the other half owns this stage
            mbar_wait_s(sm.full + my_slot * 8, my_parity, 4);
            if (!skip_math) {
                const uint32_t wb = sm.ring + my_slot * (uint32_t)a.slot_bytes + w_off;
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

// Distributed combine of the split-KV partials of one (row, kv head): the n_split CTAs that produced the partials each reduce
// a slice of the group*128 output values over ALL splits (one warp per 4 values, lanes over splits), instead of the last-arriving
// CTA reducing everything serially (that serial tail was ~5 us per layer on the critical path of every CTA).
__device__ __forceinline__ void attn_merge_slice(const DecodeParams& p, int b, int kvh, int split, int tid) {
    const int warp = tid >> 5, lane = tid & 31;
    const int group = p.nh / p.nkv;
    const int total = group * DA_HD;                                  // values of this kv head's query heads
    int per = (total + p.n_split - 1) / p.n_split;
    per = (per + 3) & ~3;
    const int e_lo = split * per, e_hi = (e_lo + per) < total ? (e_lo + per) : total;
    for (int e0 = e_lo + warp * 4; e0 < e_hi; e0 += (MK_STHREADS / 32) * 4) {
        const int head = kvh * group + e0 / DA_HD, dim = e0 % DA_HD;
        const size_t base = ((size_t)b * p.nh + head) * p.n_split;
        float M = -INFINITY;
        float mj[3], lj[3];
        float4 oj[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {                                  // n_split <= 96 here (launch checks <= 64)
            int j = lane + 32 * k;
            bool ok = j < p.n_split;
            mj[k] = ok ? __ldcg(p.part_ml + (base + j) * 2) : -INFINITY;
            lj[k] = ok ? __ldcg(p.part_ml + (base + j) * 2 + 1) : 0.f;
            oj[k] = ok ? __ldcg(reinterpret_cast<const float4*>(p.part_o + (base + j) * DA_HD + dim)) : make_float4(0.f, 0.f, 0.f, 0.f);
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
            uint2 pk;
            pk.x = f2_to_bf2x(o0 * inv, o1 * inv);
            pk.y = f2_to_bf2x(o2 * inv, o3 * inv);
            *reinterpret_cast<uint2*>(p.act + (size_t)b * p.nh * DA_HD + (size_t)head * DA_HD + dim) = pk;
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
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = gridDim.x, cta = blockIdx.x;

    if (tid == 0) {
        s_prog = 0; s_prod_wait = 0;
        for (int i = 0; i < a.n_slots; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], MK_CW / 2); }   // half the warps consume a stage
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
    unsigned int target = 0;
    uint32_t it = 0, grp = 0;
    if (warp == MK_CW)
        for (int k = 0; k < 2; ++k) {
            int eb = k * 4 + (lane >> 3);
            am.v[k][lane] = -INFINITY; am.i[k][lane] = 0x7fffffff;
            am.lo[k][lane] = eb < p.B ? p.dev_state[st_vlo(eb)] : 0;
            am.hi[k][lane] = eb < p.B ? p.dev_state[st_vhi(eb)] : 0;
        }
    // cycle accounting (PADT_MEGA_TIMELINE & 1): per phase kind [qkv | attention | o_proj | gate-up | down | logits]
    long long t_bar[6] = {0, 0, 0, 0, 0, 0}, t_work[6] = {0, 0, 0, 0, 0, 0}, t_stage[6] = {0, 0, 0, 0, 0, 0};
    long long t0 = TL ? clock64() : 0, tt = 0;
#define TSTART() do { if (TL) tt = clock64(); } while (0)
#define TADD(acc) do { if (TL) acc += clock64() - tt; } while (0)

    if (warp == 0 && lane < (int)(sizeof(DecodeLayerW) / 4)) reinterpret_cast<uint32_t*>(&s_w[0])[lane] = reinterpret_cast<const uint32_t*>(&a.layers[0])[lane];
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
    grid_sync(a.gbar, target, G, tid, a.dbg);
    const long long t_sample = TL ? clock64() - t0 : 0;

    const int qkv_n = (p.nh + 2 * p.nkv) * p.hd;
    const size_t layer_kv = (size_t)p.B * p.nkv * p.s_max * p.hd;
    uint8_t* attn_smem = xa + a.attn_off;
    // this CTA's attention item (n_split * nkv * B <= grid size: at most one item per CTA)
    const bool has_item = cta < p.n_split * p.nkv * p.B && !(a.dbg & 8);
    const int a_split = cta % p.n_split, a_kvh = (cta / p.n_split) % p.nkv, a_b = cta / (p.n_split * p.nkv);
    auto attn_args = [&](int li_) {          // rebuilt where needed: kernel parameters live in constant memory, registers are scarce
        DecAttnArgs da;
        da.qkv = p.qkv; da.cos_tab = p.cos_tab; da.sin_tab = p.sin_tab; da.max_pos = p.max_pos; da.sec0 = 0; da.sec1 = 0;
        da.state = p.dev_state; da.part_o = p.part_o; da.part_ml = p.part_ml; da.counters = p.attn_counters;
        da.out = p.act; da.B = p.B; da.nh = p.nh; da.nkv = p.nkv; da.s_max = p.s_max; da.n_split = p.n_split;
        da.scale_log2 = rsqrtf((float)p.hd) * 1.4426950408889634f;
        da.kcache = p.kcache + li_ * layer_kv; da.vcache = p.vcache + li_ * layer_kv;
        return da;
    };
    int a_kb = 0, a_ke = 0, a_cur = 0;
    if (has_item) decode_attn_range(attn_args(0), a_split, a_b, a_kb, a_ke, a_cur);   // state[0] was advanced by the sample phase of THIS launch
    const unsigned int step_idx = (unsigned int)(p.dev_state[0] - 1);       // decode launches since the prefill
    int plo = 0, phi = 0;
    proto_range(p, plo, phi);

    // ---- ONE phase loop (a single copy of the staging / consume / barrier code): 5 phases per layer, then head rows, then
    //      the prototype rows of the request
    const int n_phase = p.n_layers * 5 + 2;
    bool kv_in_flight = false;
    uint32_t next_w = 0;
    static_assert(sizeof(DecodeLayerW) % 4 == 0 && sizeof(DecodeLayerW) / 4 <= 32, "layer descriptor is copied by one warp");
    for (int ph = 0; ph < n_phase; ++ph) {
        const int li = ph / 5, k = ph < p.n_layers * 5 ? ph % 5 : 5 + (ph - p.n_layers * 5);   // k: 0..4 layer phases, 5 head, 6 prototypes
#ifdef PADT_MK_WATCHDOG
        if (tid == 0) { WD_MARK(g_wd_phase, ph); WD_MARK(g_wd_sub, 0); }
#endif
        // next layer's descriptor: the global load is issued at the start of the layer and parked in a register, the shared-memory store
        // happens three phases later.  Storing right after the load made warp 0 -- and with it every thread at the staging barrier -- wait
        // for an L2/HBM round trip once per layer (same-box A/B, profiles/r02/call16: 1.720 -> 1.705 ms per token)
        if (warp == 0 && lane < (int)(sizeof(DecodeLayerW) / 4) && li + 1 < p.n_layers) {
            if (k == 0) next_w = reinterpret_cast<const uint32_t*>(&a.layers[li + 1])[lane];
            if (k == 3) reinterpret_cast<uint32_t*>(&s_w[(li + 1) & 1])[lane] = next_w;
        }
        const DecodeLayerW& w = s_w[li & 1];       // fields are read at their point of use (keeps 22 registers free)
        if (k == 0) {
            kv_in_flight = false;
            if (has_item && a_ke > a_kb && (a.flags & MF_KV_SMEM)) {
                // the cache rows of earlier steps do not depend on this step: pull the item's first K/V tile into shared memory
                // now, so that it is resident when the barrier after the q/k/v projection opens
                decode_attn_issue_tile(attn_args(li), a_kvh, a_b, a_kb, a_ke, a_cur, attn_smem, tid, MK_STHREADS);
                kv_in_flight = true;
            }
        }
        if (k == 1) {
            // ---- M-RoPE + KV append + split-KV attention; the n_split CTAs of a (row, kv head) then combine slices of the result
            TSTART();
            if (has_item) {
                if (tid == 0) WD_MARK(g_wd_sub, 3);
                decode_attn_body<false>(attn_args(li), a_split, a_kvh, a_b, attn_smem, tid, MK_STHREADS, CSync(), kv_in_flight);
                CSync()();      // orders the partial stores of all warps before thread 0's gpu-scope release below (as in grid_sync)
                if (tid == 0) WD_MARK(g_wd_sub, 4);
                if (tid == 0) {
                    unsigned int* cnt = reinterpret_cast<unsigned int*>(p.attn_counters) + p.B * p.nkv + a_b * p.nkv + a_kvh;   // monotonic counters
                    const unsigned int want = (step_idx * (unsigned int)p.n_layers + (unsigned int)li + 1u) * (unsigned int)p.n_split;
                    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(cnt) : "memory");
#ifdef PADT_MK_WATCHDOG
                    WD_DECL(7);
                    while (ld_acquire_u32(cnt) < want) WD_CHECK();
#else
                    while (ld_acquire_u32(cnt) < want) {}
#endif
                }
                CSync()();
                if (tid == 0) WD_MARK(g_wd_sub, 5);
                attn_merge_slice(p, a_b, a_kvh, a_split, tid);
                if (tid == 0) WD_MARK(g_wd_sub, 6);
            }
            TADD(t_work[1]); TSTART();
            grid_sync(a.gbar, target, G, tid, a.dbg);
            TADD(t_bar[1]);
            continue;
        }
        // ---- projection phases
        Seg s;
        const bf16 *src, *ln = nullptr, *bias = nullptr;
        bf16 *y = nullptr, *copy0 = nullptr, *copy1 = nullptr;
        int K, ep, n_out = 0;
        bool stage = true;
        switch (k) {
            case 0: s = layer_seg(a, w, SEG_QKV, cta, G); src = p.h; ln = w.ln1; K = p.D; ep = EP_BIAS; bias = w.bqkv; y = p.qkv; n_out = qkv_n; break;
            case 2: s = layer_seg(a, w, SEG_O, cta, G); src = p.act; K = p.nh * p.hd; ep = EP_RESID; y = p.h; n_out = p.D; break;
            case 3: s = layer_seg(a, w, SEG_GU, cta, G); src = p.h; ln = w.ln2; K = p.D; ep = EP_SWIGLU; y = p.act; n_out = p.I; break;
            case 4: s = layer_seg(a, w, SEG_DOWN, cta, G); src = p.act; K = p.I; ep = EP_RESID; y = p.h; n_out = p.D; break;
            case 5:
                s = make_seg(p.head, p.head_img, p.nseg_d, 0, p.vocab, p.D, cta, G, a.slot_bytes);
                src = p.h; ln = p.final_norm; K = p.D; ep = EP_LOGITS;
                if (cta == 0) {
                    copy0 = p.last_hidden;
                    const int t = p.dev_state[0];
                    if (p.hidden_steps && t < p.max_steps) copy1 = p.hidden_steps + (size_t)t * p.B * p.D;
                }
                break;
            default:
                if (phi <= plo) continue;
                s = make_seg(p.protos + (size_t)plo * p.D, nullptr, 0, p.vocab + plo, phi - plo, p.D, cta, G, a.slot_bytes);
                src = p.h; K = p.D; ep = EP_LOGITS; stage = false;      // the normalised rows are still staged
                break;
        }
        const int tk = k < 5 ? k : 5;
        TSTART();
        if (stage) {
            stage_x(src, ln, p.eps, K, p.B, xs, tid, lane, copy0, copy1, s_part, (a.flags & MF_STAGE_MIN) != 0);
            CSync()();
        }
        TADD(t_stage[tk]); TSTART();
        if (tid == 0) WD_MARK(g_wd_sub, 1);
        consume_seg(a, s, ep, sm, it, grp, tid, bias, y, n_out, am, step_idx * (unsigned int)n_phase + (unsigned int)ph + 1u);
        if (tid == 0) WD_MARK(g_wd_sub, 2);
        TADD(t_work[tk]); TSTART();
        if (k < 5) grid_sync(a.gbar, target, G, tid, a.dbg);
        TADD(t_bar[tk]);
    }
    // ---- masked arg-max partial of this CTA
    CSync()();             // entry [pass][lane]: batch row pass*4 + lane/8, weight-row lane%8
    if (tid < p.B) {
        float bv = -INFINITY; int bi = 0x7fffffff;
        for (int k = 0; k < 8; ++k) {
            float v = am.v[tid >> 2][(tid & 3) * 8 + k]; int i = am.i[tid >> 2][(tid & 3) * 8 + k];
            if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
        }
        p.amax_val[(size_t)cta * p.B + tid] = bv;
        p.amax_idx[(size_t)cta * p.B + tid] = bi;
    }
    if (cta == 0 && tid == 0) p.dev_state[3] = G;
    if (TL && (a.dbg & 1) && tid == 0 && (cta == 0 || cta == G / 2 || cta == G - 1)) {
        printf("[mega cta %d] cycles: total %lld | sample+bar %lld\n"
               "[mega cta %d] per phase (qkv attn o gu down logits): barrier %lld %lld %lld %lld %lld %lld | work %lld %lld %lld %lld %lld %lld | stage %lld %lld %lld %lld %lld %lld\n",
               cta, (long long)(clock64() - t0), t_sample, cta, t_bar[0], t_bar[1], t_bar[2], t_bar[3], t_bar[4], t_bar[5], t_work[0], t_work[1],
               t_work[2], t_work[3], t_work[4], t_work[5], t_stage[0], t_stage[1], t_stage[2], t_stage[3], t_stage[4], t_stage[5]);
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

static int decode_mega_attn_off(const DecodeParams& p) { return (p.B * (p.D + 8) * 2 + 127) / 128 * 128; }

// Plans shared memory for the persistent kernel; returns 0 and fills the outputs, or -1 when the model does not fit.
int decode_mega_plan(const DecodeParams& p, int* slot_bytes, int* n_slots, int* x_bytes, size_t* smem_bytes) {
    if (p.B > 8) return -1;
    int kmax = p.I > p.D ? p.I : p.D;
    kmax = kmax > p.nh * p.hd ? kmax : p.nh * p.hd;
    int xb = p.B * (kmax + 8) * 2;
    // the attention scratch sits behind the activation rows of the q/k/v phase: its first K/V tile is filled while they are in use
    const int attn_need = decode_mega_attn_off(p) + MK_ATTN_SMEM;
    xb = xb > attn_need ? xb : attn_need;
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
    // EVEN slot count: the two half-CTAs take alternate stages, and a half only waits on the `full` barriers of its own stages.
    // With an even ring every slot always belongs to the same half, so a waiter is never more than one phase behind its barrier.
    // With an odd ring a half skips every other use of a slot, and when it runs ahead of the other half (up to MK_NRED row groups,
    // i.e. more than a ring revolution when a group has 16 stages) the parity test passes on the OTHER half's data
    // (seen at I = 18944: 9 slots -> wrong data, double arrivals, launch failure / hang).
    ns &= ~1;
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

int decode_mega_flags() { return getenv("PADT_MEGA_FLAGS") ? atoi(getenv("PADT_MEGA_FLAGS")) : MK_DEFAULT_FLAGS; }
bool decode_mega_split() { return (decode_mega_flags() & MF_SPLIT) != 0; }

// Every precondition of the persistent kernel in ONE predicate: run_step falls back to the per-layer kernels (decode.cu) when it says no,
// instead of aborting the generation on a PADT_REQUIRE inside the launcher.
bool decode_mega_can_run(const DecodeParams& p, int n_sms) {
    int sb, ns, xb;
    size_t smem;
    if (decode_mega_plan(p, &sb, &ns, &xb, &smem)) return false;
    if (p.hd != DA_HD || p.nh % p.nkv || p.nh / p.nkv > DA_MAXG) return false;
    if (p.D % 16 || p.I % 16) return false;
    if (p.n_split > 64 || p.n_split * p.nkv * p.B > n_sms) return false;                         // one attention item per CTA
    const int qkv_groups = ((p.nh + 2 * p.nkv) * p.hd / MK_GROUP + n_sms - 1) / n_sms + 2;       // + 2: a shared group at either end
    if (qkv_groups > MK_MAXGROUPS || (p.D / MK_GROUP + n_sms - 1) / n_sms + 2 > MK_MAXGROUPS) return false;   // bias / residual preload buffer
    if (((uintptr_t)p.head & 15) || (p.protos && ((uintptr_t)p.protos & 15))) return false;     // TMA bulk copies need 16-byte aligned rows
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
    PADT_CHECK_CUDA(cudaMemsetAsync(a.gbar, 0, sizeof(unsigned int), st));
    void* args[] = {(void*)&a};
    const void* fn = a.dbg ? (const void*)decode_mega_kernel<true> : (const void*)decode_mega_kernel<false>;   // <true>: cycle accounting
    PADT_CHECK_CUDA(cudaLaunchCooperativeKernel(fn, dim3(n_sms), dim3(MK_THREADS), args, smem, st));
    ++g_padt_launches;
    return 0;
}
