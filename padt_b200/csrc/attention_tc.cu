// tcgen05 + TMA flash attention for the long-sequence attention calls: the four ViT full-attention blocks (9216 tokens, d = 80;
// HF:244-261 driven by padt.py:89-99) and the causal GQA prefill of the LLM (d = 128; HF:744-755).  Same numerics as attention.cu
// (FlashAttention-2: fp32 scores, online softmax in the exp2 domain, P rounded to bf16 before P@V, fp32 accumulation, one
// normalisation at the end) on the Blackwell tensor path:
//   * one CTA = 128 query rows of one (segment, head); K/V tiles of 128 keys stream through a 2-stage TMA ring
//     (cp.async.bulk.tensor, 128-byte swizzle; per-segment tensor maps so rows past the segment are zero-filled by the hardware);
//   * S = Q K^T: tcgen05.mma (M = 128, N = 128, K = 16) with both operands K-major in shared memory, accumulator in TMEM;
//   * the four softmax warps own one TMEM lane (= query row) per thread: row max / row sum are thread-local, S is read with
//     tcgen05.ld in 32-column blocks, P is written to shared memory in the swizzled K-major layout of an A operand;
//   * O += P V: tcgen05.mma with V as an MN-major B operand (the [keys, d] tile as TMA delivers it), accumulator in TMEM, rescaled
//     in place (tcgen05.ld / tcgen05.st) when a row's running maximum moves.
// v1 serialises S(j) -> softmax(j) -> PV(j) inside a CTA; two CTAs per SM would need < 114 KB of shared memory (it uses 192 KB).
#include <cuda.h>
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int AT_BQ = 128, AT_BK = 128, AT_THREADS = 384;   // three warpgroups: warps 0-3 / 4-7 softmax, correction, epilogue of query tile 0 / 1 (232 registers); warp 8 TMA, warp 9 MMA + TMEM owner (40 registers)
constexpr int AT_PANEL = AT_BQ * 128;                        // one 64-column panel of a 128-row tile: 16 KB
constexpr int AT_TILE = 2 * AT_PANEL;                        // a [128 x 128] bf16 tile = two panels
constexpr int AT_MAXSEG = 8;
constexpr int AT_SMEM = 2 * AT_TILE /*two Q tiles*/ + 2 * 2 * AT_TILE /*2 stages of K, V*/ + 256 /*barriers*/ + 1024 /*alignment*/;

struct AtMaps { CUtensorMap q[AT_MAXSEG], k[AT_MAXSEG], v[AT_MAXSEG]; };

__device__ __forceinline__ void at_mbar_init(uint64_t* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void at_mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void at_mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
// bounded wait: a protocol error traps (the launch returns an error) instead of hanging the GPU
__device__ __forceinline__ void at_mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; ++spins) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(b)), "r"(parity)
            : "memory");
        if (spins > (1u << 28)) __trap();
    }
}
__device__ __forceinline__ void at_tma_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
                 : "memory");
}
// K-major operand tile, 128-byte swizzle: rows of 64 bf16 (128 B), 8-row groups 1024 B apart (SBO), LBO unused
__device__ __forceinline__ uint64_t at_desc_k(const void* smem) {
    uint64_t addr = (uint64_t)(smem_u32(smem) & 0x3FFFF) >> 4;
    return addr | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// MN-major operand tile, 128-byte swizzle (cute/atom/mma_traits_sm100.hpp, make_umma_desc<Major::MN>, B128: ((8,n),(8,k)):((1,LBO),(8,SBO))
// in 16-byte units): 64 elements contiguous along MN per 128-byte row, rows = K index, 8-row groups 1024 B apart (SBO), the next 64 MN
// elements one panel (16 KB) further (LBO)
__device__ __forceinline__ uint64_t at_desc_mn(const void* smem) {
    uint64_t addr = (uint64_t)(smem_u32(smem) & 0x3FFFF) >> 4;
    return addr | ((uint64_t)(AT_PANEL >> 4) << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void at_umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void at_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void at_tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void at_tmem_st32(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]),
        "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]),
        "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void at_umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void at_tmem_ld32_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void at_tmem_st32_nowait(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]),
        "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]),
        "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ float at_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// One CTA = two 128-row query tiles of one (segment, head), ping-ponged against the tensor core: while the softmax warps of tile 0 turn
// S0(j+1) into P0(j+1) the tensor core runs P1(j) V(j) and S1(j+1), and the reverse.  TMEM (512 columns): S0 | S1 | O0 | O1, 128 fp32
// columns each; P_t (bf16, two per column) overwrites the first 64 columns of S_t once S_t is in registers and is the TMEM A operand of
// the P V mma.  Issue order of the mma thread: S0(0) S1(0) { PV0(j) S0(j+1) PV1(j) S1(j+1) }: tcgen05.mma executes in issue order, so
// S_t(j+1) cannot overwrite P_t(j) before PV_t(j) has read it, and the commit after S_t(j+1) also covers PV_t(j): when the softmax
// warps see s_ready_t they may rescale O_t.
template <int HD>
__global__ void __launch_bounds__(AT_THREADS, 1) flash_attn_tc_kernel(const __grid_constant__ AtMaps maps, AttnParams p) {
    constexpr int KT = HD / 16;                 // k-steps of Q K^T (d = 80: 5; panel 1 holds columns 64..79 + hardware zero fill)
    constexpr int OC = HD;                      // fp32 accumulator columns of O (N of the PV mma: 80 or 128, both multiples of 16)
    extern __shared__ uint8_t at_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(at_raw) + 1023) & ~uintptr_t(1023));   // 128-byte swizzle atoms: 1024-byte aligned
    uint8_t* sQ = smem;                         // two query tiles
    uint8_t* sK = sQ + 2 * AT_TILE;             // two stages
    uint8_t* sV = sK + 2 * AT_TILE;             // two stages
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * AT_TILE);
    uint64_t* q_full = bars;                    // [1]
    uint64_t* k_full = bars + 1;                // [2]
    uint64_t* k_empty = bars + 3;               // [2]
    uint64_t* v_full = bars + 5;                // [2]
    uint64_t* v_empty = bars + 7;               // [2]
    uint64_t* s_ready = bars + 9;               // [2] per query tile
    uint64_t* p_ready = bars + 11;              // [2] per query tile, 128 arrivals
    uint64_t* o_done = bars + 13;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int seg = blockIdx.z, h = blockIdx.y;
    const int qp = p.causal ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x;   // pair of query tiles; causal: the long ones first
    const int ql = p.q_len[seg];
    if (qp * 2 * AT_BQ >= ql) return;            // whole CTA leaves before any barrier / TMEM use
    const int kl = p.k_len[seg];
    const int ks = p.k_start ? p.k_start[seg] : 0;
    const int kvh = h / p.group;
    const int diag = kl - ql;                    // causal: key j visible to query i iff j <= i + diag
    int kv_end = kl;
    if (p.causal) {
        const int e = qp * 2 * AT_BQ + 2 * AT_BQ - 1 + diag + 1;
        kv_end = e < kl ? e : kl;
    }
    const int t_begin = ks / AT_BK, t_end = (kv_end + AT_BK - 1) / AT_BK;
    const int n_tiles = t_end > t_begin ? t_end - t_begin : 0;

    if (threadIdx.x == 0) {
        at_mbar_init(q_full, 1);
        for (int i = 0; i < 2; ++i) {
            at_mbar_init(&k_full[i], 1); at_mbar_init(&k_empty[i], 1); at_mbar_init(&v_full[i], 1); at_mbar_init(&v_empty[i], 1);
            at_mbar_init(&s_ready[i], 1); at_mbar_init(&p_ready[i], 128);
        }
        at_mbar_init(o_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.q[seg]) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.k[seg]) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.v[seg]) : "memory");
    }
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    // register split: the launch gives every warp 168 registers (3 warps per scheduler); 2 x 216 + 72 = 3 x 168
    if (warp >= 8) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
      if (warp == 8) {
        if (lane == 0 && n_tiles > 0) {   // ---- TMA producer: both query tiles once, then K and V tiles through two 2-stage rings
            at_mbar_expect_tx(q_full, 2 * AT_TILE);
            for (int t = 0; t < 2; ++t) {
                at_tma_3d(sQ + t * AT_TILE, &maps.q[seg], 0, (qp * 2 + t) * AT_BQ, h, q_full);
                at_tma_3d(sQ + t * AT_TILE + AT_PANEL, &maps.q[seg], 64, (qp * 2 + t) * AT_BQ, h, q_full);
            }
            for (int j = 0; j < n_tiles; ++j) {
                const int s = j & 1, row = (t_begin + j) * AT_BK;
                const uint32_t par = ((j >> 1) & 1) ^ 1;
                at_mbar_wait(&k_empty[s], par);
                at_mbar_expect_tx(&k_full[s], AT_TILE);
                at_tma_3d(sK + s * AT_TILE, &maps.k[seg], 0, row, kvh, &k_full[s]);
                at_tma_3d(sK + s * AT_TILE + AT_PANEL, &maps.k[seg], 64, row, kvh, &k_full[s]);
                at_mbar_wait(&v_empty[s], par);
                at_mbar_expect_tx(&v_full[s], AT_TILE);
                at_tma_3d(sV + s * AT_TILE, &maps.v[seg], 0, row, kvh, &v_full[s]);
                at_tma_3d(sV + s * AT_TILE + AT_PANEL, &maps.v[seg], 64, row, kvh, &v_full[s]);
            }
        }
      } else if (warp == 9) {
        if (lane == 0 && n_tiles > 0) {   // ---- MMA issuer
            // instruction descriptors: D = f32, A = B = bf16; S: both K-major, N = 128; PV: A from TMEM, B MN-major (bit 16), N = d
            const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(AT_BK >> 3) << 17) | ((uint32_t)(AT_BQ >> 4) << 24);
            const uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(OC >> 3) << 17) | ((uint32_t)(AT_BQ >> 4) << 24);
            auto issue_s = [&](int t, int st) {
                const uint64_t q0 = at_desc_k(sQ + t * AT_TILE), q1 = at_desc_k(sQ + t * AT_TILE + AT_PANEL);
                const uint64_t k0 = at_desc_k(sK + st * AT_TILE), k1 = at_desc_k(sK + st * AT_TILE + AT_PANEL);
#pragma unroll
                for (int k = 0; k < KT; ++k)       // 16 elements = 32 bytes inside the swizzled row: +2 in the address field; panel 1 from k = 4
                    at_umma(tmem_base + t * 128, (k < 4 ? q0 : q1) + 2 * (k & 3), (k < 4 ? k0 : k1) + 2 * (k & 3), idesc_s, k ? 1u : 0u);
                at_commit(&s_ready[t]);
            };
            auto issue_pv = [&](int t, int st, int j) {
                const uint64_t vdesc = at_desc_mn(sV + st * AT_TILE);
#pragma unroll
                for (int k = 0; k < AT_BK / 16; ++k)   // 16 keys: P advances 8 TMEM columns (two bf16 per column), V 16 rows = 2048 bytes
                    at_umma_ts(tmem_base + 256 + t * 128, tmem_base + t * 128 + k * 8, vdesc + (uint64_t)(k * 128), idesc_o, (j | k) ? 1u : 0u);
            };
            at_mbar_wait(q_full, 0);
            at_mbar_wait(&k_full[0], 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            issue_s(0, 0);
            issue_s(1, 0);
            at_commit(&k_empty[0]);
            for (int j = 0; j < n_tiles; ++j) {
                const int s = j & 1, sn = s ^ 1;
                const bool more = j + 1 < n_tiles;
                at_mbar_wait(&v_full[s], (j >> 1) & 1);
                at_mbar_wait(&p_ready[0], j & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                issue_pv(0, s, j);
                if (more) {
                    at_mbar_wait(&k_full[sn], ((j + 1) >> 1) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    issue_s(0, sn);
                }
                at_mbar_wait(&p_ready[1], j & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                issue_pv(1, s, j);
                at_commit(&v_empty[s]);
                if (more) {
                    issue_s(1, sn);
                    at_commit(&k_empty[sn]);
                }
            }
            at_commit(o_done);
        }
      }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
        // ---- softmax / correction / epilogue: warps 0-3 own query tile 0, warps 4-7 tile 1; warp w touches TMEM lanes [32 (w % 4), + 32);
        // one thread = one query row
        const int t = warp >> 2;
        const int qd = warp & 3;
        const int r = qd * 32 + lane;                              // row inside the tile
        const int i0 = (qp * 2 + t) * AT_BQ;                       // first query row of the tile
        const int i = i0 + r;
        const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
        const uint32_t tmem_s = tmem_base + t * 128 + lane_addr, tmem_o = tmem_base + 256 + t * 128 + lane_addr;
        const float sl2 = p.scale * 1.4426950408889634f;
        float m_run = -INFINITY, l_run = 0.f;
        for (int j = 0; j < n_tiles; ++j) {
            const int key0 = (t_begin + j) * AT_BK;
            at_mbar_wait(&s_ready[t], j & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t sv[128];
#pragma unroll
            for (int cb = 0; cb < 128; cb += 32) at_tmem_ld32_nowait(tmem_s + (uint32_t)cb, sv + cb);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            // interior tiles carry no mask at all (warp-uniform test on the tile's corners)
            const bool need_mask = key0 + AT_BK > kl || key0 < ks || (p.causal && key0 + AT_BK - 1 > i0 + diag);
            if (need_mask) {
                const int hi = p.causal ? (i + diag < kl - 1 ? i + diag : kl - 1) : kl - 1;   // last visible key of this row
#pragma unroll
                for (int c = 0; c < 128; ++c) {
                    const int kj = key0 + c;
                    if (kj > hi || kj < ks) sv[c] = 0xff800000u;
                }
            }
            float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
            for (int c = 0; c < 128; c += 4) {
                mx0 = fmaxf(mx0, __uint_as_float(sv[c])); mx1 = fmaxf(mx1, __uint_as_float(sv[c + 1]));
                mx2 = fmaxf(mx2, __uint_as_float(sv[c + 2])); mx3 = fmaxf(mx3, __uint_as_float(sv[c + 3]));
            }
            const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * sl2;
            const float m_new = fmaxf(m_run, mx);
            const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
            const float alpha = exp2f(m_run - m_use);
            m_run = m_new;
            // P = exp2(s * scale*log2e - m) in fp32, row sum in fp32, P rounded to bf16 into TMEM over S (two keys per column)
            float rs0 = 0.f, rs1 = 0.f;
            const float neg_m = -m_use;
#pragma unroll
            for (int c = 0; c < 128; c += 2) {
                const float p0 = at_ex2(fmaf(__uint_as_float(sv[c]), sl2, neg_m));
                const float p1 = at_ex2(fmaf(__uint_as_float(sv[c + 1]), sl2, neg_m));
                rs0 += p0; rs1 += p1;
                sv[c >> 1] = f2_to_bf2x(p0, p1);
            }
            at_tmem_st32_nowait(tmem_s, sv);
            at_tmem_st32_nowait(tmem_s + 32u, sv + 32);
            // (after P: the 128 score registers are dead by now)  s_ready(j) also covers P V of tile j - 1 (committed after it): O may be rescaled now
            if (j > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll
                for (int cb = 0; cb < OC; cb += 32) {
                    uint32_t ov[32];
                    if (cb + 32 <= OC) {
                        at_tmem_ld32(tmem_o + (uint32_t)cb, ov);
#pragma unroll
                        for (int c = 0; c < 32; ++c) ov[c] = __float_as_uint(__uint_as_float(ov[c]) * alpha);
                        at_tmem_st32_nowait(tmem_o + (uint32_t)cb, ov);
                    } else {   // d = 80: columns 48..79, of which 48..63 are already scaled
                        at_tmem_ld32(tmem_o + (uint32_t)(OC - 32), ov);
#pragma unroll
                        for (int c = 16; c < 32; ++c) ov[c] = __float_as_uint(__uint_as_float(ov[c]) * alpha);
                        at_tmem_st32_nowait(tmem_o + (uint32_t)(OC - 32), ov);
                    }
                }
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            at_mbar_arrive(&p_ready[t]);
            l_run = l_run * alpha + (rs0 + rs1);
        }
        // ---- normalise + store
        bf16* ob = p.o + (long long)(p.q_off[seg] + i) * p.o_row + (long long)h * p.o_head;
        if (n_tiles > 0) {
            at_mbar_wait(o_done, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;
#pragma unroll
        for (int cb = 0; cb < OC; cb += 32) {
            const int c0 = cb + 32 <= OC ? cb : OC - 32;                      // d = 80: the last block re-reads columns 48..79
            uint32_t ov[32];
            if (n_tiles > 0) at_tmem_ld32(tmem_o + (uint32_t)c0, ov);
            if (i < ql) {
#pragma unroll
                for (int c = (cb - c0); c < 32; c += 8) {
                    uint32_t w[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        w[e] = n_tiles > 0 ? f2_to_bf2x(__uint_as_float(ov[c + 2 * e]) * inv, __uint_as_float(ov[c + 2 * e + 1]) * inv) : 0u;
                    *reinterpret_cast<uint4*>(ob + c0 + c) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 9) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn at_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// 3-D view [heads, rows, d] of one segment inside a packed buffer (element strides `head` and `row`, d contiguous), boxes of 64 columns x
// 128 rows of one head, 128-byte swizzle.  Rows past the segment and columns past d are zero-filled by the hardware, so a tile never sees
// the next segment's rows or the not-yet-written decode slots of the KV cache.
int at_make_map(CUtensorMap* map, const bf16* base, int hd, long long rows, int heads, long long row, long long head) {
    EncodeTiledFn enc = at_encode();
    PADT_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t gdim[3] = {(cuuint64_t)hd, (cuuint64_t)rows, (cuuint64_t)heads};
    cuuint64_t gstride[2] = {(cuuint64_t)row * 2, (cuuint64_t)head * 2};
    cuuint32_t box[3] = {64u, (cuuint32_t)AT_BQ, 1u};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<bf16*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PADT_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (rc " + std::to_string((int)r) + ")");
    return 0;
}

template <int HD>
int launch_tc_hd(const AttnParams& p, const AttnSegHost& hs, cudaStream_t st) {
    static bool configured_dev[64] = {false};
    bool& configured = configured_dev[padt_dev_slot()];
    if (!configured) {
        PADT_CHECK_CUDA(cudaFuncSetAttribute(flash_attn_tc_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
        configured = true;
    }
    AtMaps maps;
    const int nkv = p.n_heads / p.group;
    for (int s = 0; s < p.n_seg; ++s) {
        int rc = at_make_map(&maps.q[s], p.q + (long long)hs.q_off[s] * p.q_row, HD, hs.q_len[s], p.n_heads, p.q_row, p.q_head);
        if (rc) return rc;
        rc = at_make_map(&maps.k[s], p.k + (long long)hs.k_off[s] * p.k_row, HD, hs.k_len[s], nkv, p.k_row, p.k_head);
        if (rc) return rc;
        rc = at_make_map(&maps.v[s], p.v + (long long)hs.k_off[s] * p.v_row, HD, hs.k_len[s], nkv, p.v_row, p.v_head);
        if (rc) return rc;
    }
    dim3 grid(cdiv(p.max_q_len, 2 * AT_BQ), p.n_heads, p.n_seg);
    flash_attn_tc_kernel<HD><<<grid, AT_THREADS, AT_SMEM, st>>>(maps, p);
    PADT_LAUNCH_CHECK();
    return 0;
}

}  // namespace

// Shapes the tcgen05 attention takes: long query segments (>= 512 rows: the ViT full-attention blocks and the LLM prefill), at most 8
// segments, 16-byte aligned strides, host copies of the segment table (the tensor maps are built per segment).
bool attention_tc_eligible(const AttnParams& p, const AttnSegHost* hs) {
    static const int enabled = getenv("PADT_ATTN_TC") ? atoi(getenv("PADT_ATTN_TC")) : 1;
    if (!enabled || !hs || p.n_split > 1) return false;
    if (p.head_dim != 80 && p.head_dim != 128) return false;
    if (p.n_seg < 1 || p.n_seg > AT_MAXSEG || p.max_q_len < 512) return false;
    if (p.q_row % 8 || p.k_row % 8 || p.v_row % 8 || p.q_head % 8 || p.k_head % 8 || p.v_head % 8 || p.o_row % 8 || p.o_head % 8) return false;
    if ((((uintptr_t)p.q) | ((uintptr_t)p.k) | ((uintptr_t)p.v) | ((uintptr_t)p.o)) & 15) return false;
    return at_encode() != nullptr;
}

int launch_attention_tc(const AttnParams& p, const AttnSegHost& hs, cudaStream_t st) {
    if (p.head_dim == 80) return launch_tc_hd<80>(p, hs, st);
    return launch_tc_hd<128>(p, hs, st);
}
