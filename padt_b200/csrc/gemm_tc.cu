// tcgen05 + TMA dense projection for the large-M stages (ViT blocks / merger, LLM prefill, PaDT-decoder memory side):
//   C[M,N] = A[M,K] * W[N,K]^T (+bias) with the same fused epilogues and HF rounding points as gemm.cu.
//
// One CTA computes a 128 x BN (BN = 256 or 128) tile: a TMA producer thread streams 64-column K panels of A and W
// (cp.async.bulk.tensor, 128-byte swizzle) through a 4-stage shared-memory ring, one elected thread issues
// tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16) into a TMEM accumulator (BN fp32 columns), tcgen05.commit
// releases the ring slots, and four epilogue warps read the accumulator back with tcgen05.ld (one TMEM lane = one output
// row per thread), apply bias / GELU / residual / interleaved SwiGLU and store bf16 rows.
// Replaces the cuBLAS calls behind every nn.Linear of the reference's tensor-core-bound stages.
#include <cuda.h>
#include <cstdlib>
#include <map>
#include <mutex>
#include <utility>
#include "common.cuh"
#include "kernels.h"

#define RC_RET(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)

namespace {

constexpr int TC_BM = 128, TC_BK = 64, TC_STAGES = 4;
constexpr int TC_THREADS = 192;   // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-5: epilogue
constexpr int TC_THREADS_P = 320; // persistent kernel: warps 2-9 = epilogue (two warps per TMEM lane quarter, half of the columns each)

// SiLU of the SwiGLU epilogue with the fast exponential and division (ex2.approx / rcp.approx based, a few fp32 ulps): the SwiGLU
// epilogue is math-bound (two MUFU + ~25 instructions per pair with expf and IEEE division: 24 us of the ViT gate/up GEMM), and the
// value is rounded to bf16 (2^-9 relative) right afterwards -- the rounding changes for about 1 element in 10^4.
__device__ __forceinline__ float silu_fast_f(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

__device__ __forceinline__ void tc_mbar_init(uint64_t* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void tc_mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint64_t* b, uint32_t parity) {
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
// bounded variant (persistent kernels): a protocol error traps -- the launch fails -- instead of hanging the GPU
__device__ __forceinline__ void tp_or_tc_wait(uint64_t* b, uint32_t parity) {
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
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
// K-major operand tile, 128-byte swizzle: rows of 64 bf16 (128 B), 8-row groups 1024 B apart (SBO), LBO unused (= 1)
__device__ __forceinline__ uint64_t umma_desc_k_sw128(const void* smem) {
    uint64_t addr = (uint64_t)(smem_u32(smem) & 0x3FFFF) >> 4;
    return addr | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t* r) {
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

// Epilogue of one 32-column block of one output row (r = fp32 accumulators read from TMEM): bias / GELU / residual /
// interleaved SwiGLU with the HF rounding points, 16-byte stores.
template <int EPI>
__device__ __forceinline__ void epilogue_block(const GemmParams& p, const uint32_t* r, int orow, int col0) {
    const bool full = col0 + 32 <= p.N;
    float v[32];
    // bias: the 32 columns of this block are the same for every lane -> four broadcast 16-byte loads
    if (p.bias) {
        const bf16* bp = p.bias + col0;
        if (full && (reinterpret_cast<uintptr_t>(bp) & 15) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                uint4 bv = *reinterpret_cast<const uint4*>(bp + j);
                float2 b0 = bf2x_to_f2(bv.x), b1 = bf2x_to_f2(bv.y), b2 = bf2x_to_f2(bv.z), b3 = bf2x_to_f2(bv.w);
                v[j] = b0.x; v[j + 1] = b0.y; v[j + 2] = b1.x; v[j + 3] = b1.y; v[j + 4] = b2.x; v[j + 5] = b2.y; v[j + 6] = b3.x; v[j + 7] = b3.y;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = col0 + j < p.N ? bf2f(bp[j]) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = rbf(__uint_as_float(r[j]) + v[j]);
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = rbf(__uint_as_float(r[j]));
    }
    if (EPI == EPI_SWIGLU) {
        // interleaved (gate, up) column pairs -> 16 outputs = 32 contiguous bytes per thread: two 16-byte stores
        // (scalar 2-byte stores from 32 different rows cost 16x the write sectors and dominated this kernel)
        bf16* dst = p.C + (size_t)orow * p.ldc + (col0 >> 1);
        uint32_t o[8];
#pragma unroll
        for (int j = 0; j < 32; j += 4)
            o[j >> 2] = f2_to_bf2x(rbf(silu_fast_f(v[j])) * v[j + 1], rbf(silu_fast_f(v[j + 2])) * v[j + 3]);
        if (full && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
            *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<uint4*>(dst + 8) = make_uint4(o[4], o[5], o[6], o[7]);
        } else {
#pragma unroll
            for (int j = 0; j < 32; j += 2)
                if (col0 + j + 1 < p.N) dst[j >> 1] = f2bf(rbf(silu_fast_f(v[j])) * v[j + 1]);
        }
    } else {
        if (EPI == EPI_GELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = rbf(gelu_erf_f(v[j]));
        }
        if (EPI == EPI_RESID) {
            const bf16* rs = p.resid + (size_t)orow * p.ldr + col0;
            if (full && (reinterpret_cast<uintptr_t>(rs) & 15) == 0) {
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    uint4 rv = *reinterpret_cast<const uint4*>(rs + j);
                    float2 r0 = bf2x_to_f2(rv.x), r1 = bf2x_to_f2(rv.y), r2 = bf2x_to_f2(rv.z), r3 = bf2x_to_f2(rv.w);
                    v[j] = rbf(r0.x + v[j]); v[j + 1] = rbf(r0.y + v[j + 1]); v[j + 2] = rbf(r1.x + v[j + 2]); v[j + 3] = rbf(r1.y + v[j + 3]);
                    v[j + 4] = rbf(r2.x + v[j + 4]); v[j + 5] = rbf(r2.y + v[j + 5]); v[j + 6] = rbf(r3.x + v[j + 6]); v[j + 7] = rbf(r3.y + v[j + 7]);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (col0 + j < p.N) v[j] = rbf(bf2f(rs[j]) + v[j]);
            }
        }
        bf16* dst = p.C + (size_t)orow * p.ldc + col0;
        if (full && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 8)
                *reinterpret_cast<uint4*>(dst + j) = make_uint4(f2_to_bf2x(v[j], v[j + 1]), f2_to_bf2x(v[j + 2], v[j + 3]),
                                                                f2_to_bf2x(v[j + 4], v[j + 5]), f2_to_bf2x(v[j + 6], v[j + 7]));
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) dst[j] = f2bf(v[j]);
        }
    }
}

// ---- software-pipelined epilogue of the persistent kernels.  The first version waited, per 32-column block, for the TMEM load, then
// for the bias load, then (residual epilogue) for the residual load -- three dependent latencies per block with two warps per
// scheduler: ncu's source view had 40 % of all samples on the first use of the bias value, and the epilogue (10.6 us per tile at the
// ViT qkv shape), not the mainloop (8 us), set the pace: skipping it gave 20.0 instead of 24.3 ms per request.  Now the bias of the
// tile sits in shared memory (one load per tile), the next block's TMEM load and residual load are in flight while the current block
// is computed and stored.
__device__ __forceinline__ void tmem_ld_32x32b_x32_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// residual of one 32-column block of one row, as four 16-byte loads (valid only when the block is full and 16-byte aligned)
__device__ __forceinline__ bool resid_vec_ok(const GemmParams& p, int orow, int col0) {
    return col0 + 32 <= p.N && ((reinterpret_cast<uintptr_t>(p.resid + (size_t)orow * p.ldr + col0) & 15) == 0);
}
__device__ __forceinline__ void resid_prefetch(const GemmParams& p, int orow, int col0, uint4* rv) {
    const uint4* rs = reinterpret_cast<const uint4*>(p.resid + (size_t)orow * p.ldr + col0);
#pragma unroll
    for (int j = 0; j < 4; ++j) rv[j] = rs[j];
}

// same arithmetic and rounding points as epilogue_block; bias from shared memory (fp32 copies of the bf16 values, zero where there is no
// bias or past N), residual either prefetched (`rv`, when rv_ok) or read here.  Called by ALL lanes of the warp (lanes whose row is past M
// compute and drop): the bf16 results go through a per-warp shared-memory tile [32 rows][80 B] so that the global stores are whole
// 32-byte sectors of 8 (or 16) rows per instruction instead of 32 half-written sectors of 32 rows -- the row-per-lane 16-byte stores
// were 16 of the 18 us the epilogue added to the ViT qkv GEMM.
template <int EPI>
__device__ __forceinline__ void epilogue_block_sb(const GemmParams& p, const uint32_t* r, int orow, bool row_ok, int col0, const float* sb,
                                                  const uint4* rv, bool rv_ok, bf16* stage, int lane) {
    const bool full = col0 + 32 <= p.N;
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        const float4 b = *reinterpret_cast<const float4*>(sb + j);
        v[j] = rbf(__uint_as_float(r[j]) + b.x); v[j + 1] = rbf(__uint_as_float(r[j + 1]) + b.y);
        v[j + 2] = rbf(__uint_as_float(r[j + 2]) + b.z); v[j + 3] = rbf(__uint_as_float(r[j + 3]) + b.w);
    }
    const bool c_vec = full && (p.ldc % 8) == 0 && (reinterpret_cast<uintptr_t>(p.C) & 15) == 0;   // warp-uniform
    uint4* srow = reinterpret_cast<uint4*>(stage + lane * 40);
    if (EPI == EPI_SWIGLU) {
        uint32_t o[8];
#pragma unroll
        for (int j = 0; j < 32; j += 4)
            o[j >> 2] = f2_to_bf2x(rbf(silu_fast_f(v[j])) * v[j + 1], rbf(silu_fast_f(v[j + 2])) * v[j + 3]);
        if (p.dbg_skip_epilogue == 2) return;   // measurement: math without stores
        // 16 outputs = 32 contiguous bytes per row: already one whole sector per lane, no staging needed (staging made this
        // math-bound epilogue slower: more instructions per block)
        if (row_ok) {
            bf16* dst = p.C + (size_t)orow * p.ldc + (col0 >> 1);
            if (full && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
                *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
                *reinterpret_cast<uint4*>(dst + 8) = make_uint4(o[4], o[5], o[6], o[7]);
            } else {
#pragma unroll
                for (int j = 0; j < 32; j += 2)
                    if (col0 + j + 1 < p.N) dst[j >> 1] = f2bf(rbf(silu_fast_f(v[j])) * v[j + 1]);
            }
        }
        return;
    }
    if (EPI == EPI_GELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = rbf(gelu_erf_f(v[j]));
    }
    if (EPI == EPI_RESID) {
        if (rv_ok) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 r0 = bf2x_to_f2(rv[j].x), r1 = bf2x_to_f2(rv[j].y), r2 = bf2x_to_f2(rv[j].z), r3 = bf2x_to_f2(rv[j].w);
                float* w = v + 8 * j;
                w[0] = rbf(r0.x + w[0]); w[1] = rbf(r0.y + w[1]); w[2] = rbf(r1.x + w[2]); w[3] = rbf(r1.y + w[3]);
                w[4] = rbf(r2.x + w[4]); w[5] = rbf(r2.y + w[5]); w[6] = rbf(r3.x + w[6]); w[7] = rbf(r3.y + w[7]);
            }
        } else if (row_ok) {
            const bf16* rs = p.resid + (size_t)orow * p.ldr + col0;
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) v[j] = rbf(bf2f(rs[j]) + v[j]);
        }
    }
    if (p.dbg_skip_epilogue == 2) { if (row_ok && v[0] + v[31] == 1.2345e-30f) p.C[(size_t)orow * p.ldc + col0] = f2bf(v[7]); return; }   // measurement: math without stores
    if (c_vec) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            srow[j] = make_uint4(f2_to_bf2x(v[8 * j], v[8 * j + 1]), f2_to_bf2x(v[8 * j + 2], v[8 * j + 3]), f2_to_bf2x(v[8 * j + 4], v[8 * j + 5]),
                                 f2_to_bf2x(v[8 * j + 6], v[8 * j + 7]));
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {          // 8 rows x 64 bytes per instruction
            const int rr = (lane >> 2) + 8 * i, ch = lane & 3;
            const uint4 val = *reinterpret_cast<const uint4*>(stage + rr * 40 + ch * 8);
            const int orr = __shfl_sync(0xffffffffu, orow, rr);
            const int okr = __shfl_sync(0xffffffffu, (int)row_ok, rr);
            if (okr) *reinterpret_cast<uint4*>(p.C + (size_t)orr * p.ldc + col0 + ch * 8) = val;
        }
        __syncwarp();
    } else if (row_ok) {
        bf16* dst = p.C + (size_t)orow * p.ldc + col0;
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (col0 + j < p.N) dst[j] = f2bf(v[j]);
    }
}

// One tile's epilogue for the eight epilogue warps of a CTA (threads 64..319): TMEM lanes [32 q, +32), columns [c_lo, c_hi) of an
// accumulator of `bn` columns at `tmem_acc`.  `release(…)` is called once every column of the tile is in registers.
template <int EPI, typename Release>
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, uint32_t tmem_acc, int q, int lane, int c_lo, int c_hi, int m0, int n0, int bn,
                                              float* sbias, bf16* stage, uint64_t* tmem_full_bar, uint32_t full_parity, Release release) {
    const int row = m0 + q * 32 + lane;
    const bool row_ok = row < p.M;
    const int orow = row_ok ? (p.out_row_map ? p.out_row_map[row] : row) : 0;
    // bias of this tile -> shared memory (fp32), while the mainloop of the tile may still be running
    const int te = (int)threadIdx.x - 64;
    if (te < bn) sbias[te] = (p.bias && n0 + te < p.N) ? bf2f(p.bias[n0 + te]) : 0.f;
    // residual of the first block: does not depend on the accumulator either
    uint4 rv[4], rvn[4];
    bool rv_ok = false, rvn_ok = false;
    if (EPI == EPI_RESID && row_ok && n0 + c_lo < p.N) {
        rv_ok = resid_vec_ok(p, orow, n0 + c_lo);
        if (rv_ok) resid_prefetch(p, orow, n0 + c_lo, rv);
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    tp_or_tc_wait(tmem_full_bar, full_parity);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16);
    uint32_t r[32], rn[32];
    tmem_ld_32x32b_x32_nowait(taddr + (uint32_t)c_lo, r);
    tmem_ld_wait();
    for (int cb = c_lo; cb < c_hi; cb += 32) {
        const bool more = cb + 32 < c_hi;
        if (more) {
            tmem_ld_32x32b_x32_nowait(taddr + (uint32_t)(cb + 32), rn);
            if (EPI == EPI_RESID && row_ok && n0 + cb + 32 < p.N) {
                rvn_ok = resid_vec_ok(p, orow, n0 + cb + 32);
                if (rvn_ok) resid_prefetch(p, orow, n0 + cb + 32, rvn);
            }
        }
        const int col0 = n0 + cb;
        if (col0 < p.N && p.dbg_skip_epilogue != 1) epilogue_block_sb<EPI>(p, r, orow, row_ok, col0, sbias + cb, rv, rv_ok, stage, lane);
        if (more) {
            tmem_ld_wait();
            if (cb + 64 >= c_hi) release();          // the last block of this thread's columns is in registers
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = rn[j];
#pragma unroll
            for (int j = 0; j < 4; ++j) rv[j] = rvn[j];
            rv_ok = rvn_ok;
        } else if (c_hi - c_lo == 32) {
            release();
        }
    }
}

// Split-K of the last, partial round of tiles: `n_units` = (tiles past the full rounds) x `parts` CTAs each take 1 / parts of one
// tile's k-blocks and write an fp32 partial to `ws`; when all parts of a tile have arrived (counter, the CTAs of a persistent grid
// are co-resident) every part sums and finishes 1 / parts of the tile's 32-column blocks, in part order (deterministic).
struct TcSplit {
    int dp_tiles;             // tiles [0, dp_tiles) are walked whole (tile = cta, cta + grid, ...)
    int n_units, parts;       // 0 units: no split
    float* ws;                // [n_units][128][BN] fp32
    unsigned int* counters;   // [tiles past dp_tiles][2] (arrived, finished), zero between launches
};

template <int BN>
struct TcSmem {
    static constexpr int A_BYTES = TC_BM * TC_BK * 2;          // 16 KB
    static constexpr int B_BYTES = BN * TC_BK * 2;             // 32 / 16 KB
    static constexpr int STAGE = A_BYTES + B_BYTES;
    static constexpr int TOTAL = TC_STAGES * STAGE + 1024 /*alignment slack*/ + 256 /*barriers, TMEM slot*/ + 2048 /*bias of two tiles, fp32*/ + 8 * 32 * 80 /*per-warp store staging*/;
};

template <int BN, int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));   // SW128 needs 1024-byte tiles
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + TC_STAGES * TcSmem<BN>::STAGE);
    uint64_t* empty = full + TC_STAGES;
    uint64_t* tmem_full = empty + TC_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * BN;
    const int nk = (p.K + TC_BK - 1) / TC_BK;

    if (threadIdx.x == 0) {
        for (int i = 0; i < TC_STAGES; ++i) { tc_mbar_init(&full[i], 1); tc_mbar_init(&empty[i], 1); }
        tc_mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    }
    if (warp == 1) {   // TMEM: BN fp32 accumulator columns for 128 lanes
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {   // ---- TMA producer
            for (int kb = 0; kb < nk; ++kb) {
                const int s = kb % TC_STAGES;
                tc_mbar_wait(&empty[s], ((kb / TC_STAGES) & 1) ^ 1);
                tc_mbar_expect_tx(&full[s], TcSmem<BN>::STAGE);
                uint8_t* sa = smem + s * TcSmem<BN>::STAGE;
                tma_load_2d(sa, &map_a, kb * TC_BK, m0, &full[s]);
                tma_load_2d(sa + TcSmem<BN>::A_BYTES, &map_w, kb * TC_BK, n0, &full[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {   // ---- MMA issuer
            // instruction descriptor: D=f32, A=B=bf16, both K-major, N=BN, M=128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
            for (int kb = 0; kb < nk; ++kb) {
                const int s = kb % TC_STAGES;
                tc_mbar_wait(&full[s], (kb / TC_STAGES) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint8_t* sa = smem + s * TcSmem<BN>::STAGE;
                const uint64_t adesc = umma_desc_k_sw128(sa), bdesc = umma_desc_k_sw128(sa + TcSmem<BN>::A_BYTES);
#pragma unroll
                for (int k = 0; k < TC_BK / 16; ++k)   // advance 16 elements = 32 bytes inside the swizzled row: +2 in the address field
                    umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) ? 1u : 0u);
                umma_commit(&empty[s]);                // frees the slot when these MMAs have read it
            }
            umma_commit(tmem_full);                    // accumulator complete
        }
    } else {
        // ---- epilogue: warp w may only touch TMEM lanes [32*(w%4), 32*(w%4)+32); lane = output row inside the tile
        tc_mbar_wait(tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3;
        const int row = m0 + q * 32 + lane;
        const bool row_ok = row < p.M;
        const int orow = row_ok ? (p.out_row_map ? p.out_row_map[row] : row) : 0;
        for (int cb = 0; cb < BN; cb += 32) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cb, r);
            const int col0 = n0 + cb;
            if (!row_ok || col0 >= p.N) continue;
            epilogue_block<EPI>(p, r, orow, col0);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(BN) : "memory");
}

// The unpipelined tile epilogue (block by block: TMEM load, epilogue_block).  Kept for the bias-free SwiGLU projections (LLM gate/up):
// their epilogue is math-bound, has no bias / residual latency to hide, and the pipelined version's extra moves and per-tile barrier cost
// 5 % there (151.6 -> 160 us at 2350 x 22016 x 2048).
template <int EPI, typename Release>
__device__ __forceinline__ void epilogue_tile_simple(const GemmParams& p, uint32_t tmem_acc, int q, int lane, int c_lo, int c_hi, int m0, int n0,
                                                     uint64_t* tmem_full_bar, uint32_t full_parity, Release release) {
    tp_or_tc_wait(tmem_full_bar, full_parity);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = m0 + q * 32 + lane;
    const bool row_ok = row < p.M;
    const int orow = row_ok ? (p.out_row_map ? p.out_row_map[row] : row) : 0;
    for (int cb = c_lo; cb < c_hi; cb += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)cb, r);
        if (cb + 32 >= c_hi) release();
        const int col0 = n0 + cb;
        if (!row_ok || col0 >= p.N || p.dbg_skip_epilogue) continue;
        epilogue_block<EPI>(p, r, orow, col0);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Persistent variant: one CTA per SM walks tiles (tile = cta, cta + grid, ...).  The TMEM holds TWO accumulators
// (2 x BN fp32 columns), so the MMA issuer starts tile i+1 while the four epilogue warps drain tile i, and the
// shared-memory ring runs across tile boundaries (no pipeline drain / refill, no per-tile TMEM allocation, barrier
// initialisation or tensor-map fetch).  Same arithmetic as the single-tile kernel.
// ---------------------------------------------------------------------------------------------------------------
template <int BN, int EPI>
__global__ void __launch_bounds__(TC_THREADS_P, 1)
gemm_tc_persistent_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, GemmParams p, int tiles_m, int n_tiles,
                          TcSplit sp) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + TC_STAGES * TcSmem<BN>::STAGE);
    uint64_t* empty = full + TC_STAGES;
    uint64_t* tmem_full = empty + TC_STAGES;     // [2]
    uint64_t* tmem_empty = tmem_full + 2;        // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nk = (p.K + TC_BK - 1) / TC_BK;
    // split-K unit of this CTA (if any): tile sp.dp_tiles + u / parts, k-blocks [kb0, kb1) of it
    const bool has_split = (int)blockIdx.x < sp.n_units;
    const int sk_tile = sp.dp_tiles + (has_split ? (int)blockIdx.x / sp.parts : 0), sk_part = has_split ? (int)blockIdx.x % sp.parts : 0;
    const int sk_kb0 = (int)((long long)nk * sk_part / sp.parts), sk_kb1 = (int)((long long)nk * (sk_part + 1) / sp.parts);

    if (threadIdx.x == 0) {
        for (int i = 0; i < TC_STAGES; ++i) { tc_mbar_init(&full[i], 1); tc_mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { tc_mbar_init(&tmem_full[i], 1); tc_mbar_init(&tmem_empty[i], 8); }   // 8 epilogue warps release an accumulator
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(2 * BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {   // ---- TMA producer: k-blocks of all of this CTA's tiles, back to back
            uint32_t it = 0;
            auto load_range = [&](int tile, int kb_lo, int kb_hi) {
                const int m0 = (tile % tiles_m) * TC_BM, n0 = (tile / tiles_m) * BN;
                for (int kb = kb_lo; kb < kb_hi; ++kb, ++it) {
                    const int s = it % TC_STAGES;
                    tc_mbar_wait(&empty[s], ((it / TC_STAGES) & 1) ^ 1);
                    tc_mbar_expect_tx(&full[s], TcSmem<BN>::STAGE);
                    uint8_t* sa = smem + s * TcSmem<BN>::STAGE;
                    tma_load_2d(sa, &map_a, kb * TC_BK, m0, &full[s]);
                    tma_load_2d(sa + TcSmem<BN>::A_BYTES, &map_w, kb * TC_BK, n0, &full[s]);
                }
            };
            for (int tile = blockIdx.x; tile < sp.dp_tiles; tile += gridDim.x) load_range(tile, 0, nk);
            if (has_split) load_range(sk_tile, sk_kb0, sk_kb1);
        }
    } else if (warp == 1) {
        if (lane == 0) {   // ---- MMA issuer
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
            uint32_t it = 0, ti = 0;
            auto mma_range = [&](int n_kb) {
                const uint32_t acc = ti & 1;
                tc_mbar_wait(&tmem_empty[acc], ((ti >> 1) & 1) ^ 1);      // the epilogue has drained this accumulator (first two uses pass)
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_d = tmem_base + acc * BN;
                for (int kb = 0; kb < n_kb; ++kb, ++it) {
                    const int s = it % TC_STAGES;
                    tc_mbar_wait(&full[s], (it / TC_STAGES) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint8_t* sa = smem + s * TcSmem<BN>::STAGE;
                    const uint64_t adesc = umma_desc_k_sw128(sa), bdesc = umma_desc_k_sw128(sa + TcSmem<BN>::A_BYTES);
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; ++k) umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) ? 1u : 0u);
                    umma_commit(&empty[s]);
                }
                umma_commit(&tmem_full[acc]);
                ++ti;
            };
            for (int tile = blockIdx.x; tile < sp.dp_tiles; tile += gridDim.x) mma_range(nk);
            if (has_split) mma_range(sk_kb1 - sk_kb0);
        }
    } else {
        // ---- epilogue warps: TMEM lanes [32*(warp%4), +32); lane = output row inside the tile; warps 2-5 take the first half of
        //      the accumulator columns, warps 6-9 the second half (the SwiGLU / bias epilogues are otherwise slower than a K = 1280 mainloop)
        const int q = warp & 3;
        const int c_lo = warp < 6 ? 0 : BN / 2, c_hi = c_lo + BN / 2;
        uint32_t ti = 0;
        float* sbias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full) + 256);
        bf16* stage = reinterpret_cast<bf16*>(reinterpret_cast<uint8_t*>(full) + 256 + 2048) + (warp - 2) * (32 * 40);   // per-warp [32][40] bf16
        for (int tile = blockIdx.x; tile < sp.dp_tiles; tile += gridDim.x, ++ti) {
            const int m0 = (tile % tiles_m) * TC_BM, n0 = (tile / tiles_m) * BN;
            const uint32_t acc = ti & 1;
            auto release = [&]() {
                // every accumulator column of this warp is in registers: hand the accumulator back before the last block's math
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty[acc])) : "memory");
            };
            if (EPI == EPI_SWIGLU && p.bias == nullptr)
                epilogue_tile_simple<EPI>(p, tmem_base + acc * BN, q, lane, c_lo, c_hi, m0, n0, &tmem_full[acc], (ti >> 1) & 1, release);
            else
                epilogue_tile<EPI>(p, tmem_base + acc * BN, q, lane, c_lo, c_hi, m0, n0, BN, sbias + (ti & 1) * 256, stage, &tmem_full[acc], (ti >> 1) & 1, release);
        }
        if (has_split) {
            // ---- split-K unit: the fp32 partial of k-blocks [kb0, kb1) goes to this unit's workspace slot; the CTA whose arrival completes
            // the tile sums the parts in part order (deterministic) and runs the epilogue
            const int m0 = (sk_tile % tiles_m) * TC_BM, n0 = (sk_tile / tiles_m) * BN;
            const uint32_t acc = ti & 1;
            tc_mbar_wait(&tmem_full[acc], (ti >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int rt = q * 32 + lane;                              // row inside the tile
            float* slot = sp.ws + ((size_t)blockIdx.x * TC_BM + rt) * BN;
            for (int cb = c_lo; cb < c_hi; cb += 32) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + (uint32_t)cb, r);
#pragma unroll
                for (int j = 0; j < 32; j += 4) __stcg(reinterpret_cast<uint4*>(slot + cb + j), make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]));
            }
            __threadfence();
            asm volatile("bar.sync 1, 256;" ::: "memory");            // the eight epilogue warps
            unsigned int* arrive = sp.counters + 2 * (sk_tile - sp.dp_tiles);
            if (threadIdx.x == 64) {
                atomicAdd(arrive, 1u);
                unsigned int seen;
                do {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(arrive) : "memory");
                } while (seen < (unsigned int)sp.parts);
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            __threadfence();
            // every part of this tile is in the workspace: this CTA finishes the 32-column blocks part, part + parts, ... for all 128 rows
            const float* base = sp.ws + (size_t)(blockIdx.x - sk_part) * TC_BM * BN;   // part 0 of this tile
            const int tid_e = threadIdx.x - 64;
            const int n_blk = (BN / 32 - sk_part + sp.parts - 1) / sp.parts;
            for (int item = tid_e; item < TC_BM * n_blk; item += 256) {
                const int rt2 = item & (TC_BM - 1), cb = (sk_part + (item >> 7) * sp.parts) * 32;
                const int row = m0 + rt2, col0 = n0 + cb;
                if (row >= p.M || col0 >= p.N) continue;
                const int orow = p.out_row_map ? p.out_row_map[row] : row;
                const float* src = base + (size_t)rt2 * BN + cb;
                float a[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) a[j] = 0.f;
                for (int part = 0; part < sp.parts; part += 2) {       // two parts = sixteen 16-byte loads in flight per thread
                    const bool two = part + 1 < sp.parts;
                    float4 v0[8], v1[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        v0[j] = __ldcg(reinterpret_cast<const float4*>(src + (size_t)part * TC_BM * BN) + j);
                        v1[j] = two ? __ldcg(reinterpret_cast<const float4*>(src + (size_t)(part + 1) * TC_BM * BN) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {      // part order: deterministic
                        a[4 * j] += v0[j].x; a[4 * j + 1] += v0[j].y; a[4 * j + 2] += v0[j].z; a[4 * j + 3] += v0[j].w;
                        a[4 * j] += v1[j].x; a[4 * j + 1] += v1[j].y; a[4 * j + 2] += v1[j].z; a[4 * j + 3] += v1[j].w;
                    }
                }
                uint32_t r[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(a[j]);
                epilogue_block<EPI>(p, r, orow, col0);
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (threadIdx.x == 64) {
                const unsigned int old = atomicAdd(arrive + 1, 1u);
                if (old == (unsigned int)(sp.parts - 1)) { arrive[0] = 0; arrive[1] = 0; }     // every part has read: self-cleaning for the next launch
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BN) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): a cluster of two CTAs on one TPC computes a 256 x 256 tile.  Each CTA stages ITS 128 rows of A and
// ITS half (128 rows) of the W tile; one thread of the even CTA issues tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 16), which reads
// A from both CTAs and exchanges the W halves between the two SMs, and each CTA's TMEM receives its own 128 x 256 accumulator.
// Why: with cta_group::1 a 128 x 256 tile moves 12 KB per k-step through ONE SM's shared memory twice (TMA write + tensor-core read =
// 192 B/clk against 128 B/clk available), which caps the mainloop at ~2/3 of the tensor peak; the pair moves 8 KB per SM (128 B/clk).
// Protocol: `full[s]` lives in the even CTA and counts the bytes of both CTAs' TMA loads (the odd CTA's loads signal it through the
// cluster address space); tcgen05.commit.multicast releases `empty[s]` / publishes `tmem_full[a]` in both CTAs; the epilogue warps of
// both CTAs hand an accumulator back on the even CTA's `tmem_empty[a]` (16 arrivals).  Every wait is bounded (trap, not hang).
// ---------------------------------------------------------------------------------------------------------------
constexpr int TP_STAGES = 6, TP_BN = 256;
constexpr int TP_A_BYTES = TC_BM * TC_BK * 2, TP_B_BYTES = (TP_BN / 2) * TC_BK * 2, TP_STAGE = TP_A_BYTES + TP_B_BYTES;   // 16 + 16 KB
constexpr int TP_SMEM = TP_STAGES * TP_STAGE + 1024 + 256 + 2048 + 8 * 32 * 80;

__device__ __forceinline__ void tp_wait(uint64_t* b, uint32_t parity) { tp_or_tc_wait(b, parity); }
__device__ __forceinline__ uint32_t tp_map_to_cta(const void* local_smem, uint32_t cta) {   // shared::cluster address of the same offset in `cta`
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(local_smem)), "r"(cta));
    return r;
}
__device__ __forceinline__ void tp_tma_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint32_t bar_cluster_addr) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(bar_cluster_addr)
                 : "memory");
}
__device__ __forceinline__ void tp_umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tp_commit_both(uint64_t* bar) {   // arrives on `bar` (same offset) in both CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((uint16_t)3)
                 : "memory");
}

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS_P, 1)
gemm_tc_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, GemmParams p, int tiles_m, int n_tiles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + TP_STAGES * TP_STAGE);
    uint64_t* empty = full + TP_STAGES;
    uint64_t* tmem_full = empty + TP_STAGES;     // [2]
    uint64_t* tmem_empty = tmem_full + 2;        // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int nk = (p.K + TC_BK - 1) / TC_BK;

    if (threadIdx.x == 0) {
        for (int i = 0; i < TP_STAGES; ++i) { tc_mbar_init(&full[i], 1); tc_mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { tc_mbar_init(&tmem_full[i], 1); tc_mbar_init(&tmem_empty[i], 16); }   // 8 epilogue warps x 2 CTAs
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    }
    if (warp == 1) {   // both CTAs, same warp id: collective allocation of 2 x 256 columns in both SMs
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    // barriers of BOTH CTAs must be initialised before either side signals across
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {   // ---- TMA producer of THIS CTA: its 128 rows of A, its 128 rows of the W tile; bytes counted on the even CTA's barrier
            uint32_t it = 0;
            for (int tile = pair; tile < n_tiles; tile += n_pairs) {
                const int m0 = (tile % tiles_m) * 2 * TC_BM + (int)rank * TC_BM, n0 = (tile / tiles_m) * TP_BN + (int)rank * (TP_BN / 2);
                for (int kb = 0; kb < nk; ++kb, ++it) {
                    const int s = it % TP_STAGES;
                    tp_wait(&empty[s], ((it / TP_STAGES) & 1) ^ 1);
                    if (rank == 0) tc_mbar_expect_tx(&full[s], 2 * TP_STAGE);
                    const uint32_t bar = tp_map_to_cta(&full[s], 0);
                    uint8_t* sa = smem + s * TP_STAGE;
                    tp_tma_2d(sa, &map_a, kb * TC_BK, m0, bar);
                    tp_tma_2d(sa + TP_A_BYTES, &map_w, kb * TC_BK, n0, bar);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {   // ---- MMA issuer of the pair
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TP_BN >> 3) << 17) | ((uint32_t)((2 * TC_BM) >> 4) << 24);
            uint32_t it = 0, ti = 0;
            for (int tile = pair; tile < n_tiles; tile += n_pairs, ++ti) {
                const uint32_t acc = ti & 1;
                tp_wait(&tmem_empty[acc], ((ti >> 1) & 1) ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_d = tmem_base + acc * TP_BN;
                for (int kb = 0; kb < nk; ++kb, ++it) {
                    const int s = it % TP_STAGES;
                    tp_wait(&full[s], (it / TP_STAGES) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint8_t* sa = smem + s * TP_STAGE;
                    const uint64_t adesc = umma_desc_k_sw128(sa), bdesc = umma_desc_k_sw128(sa + TP_A_BYTES);
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; ++k) tp_umma(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) ? 1u : 0u);
                    tp_commit_both(&empty[s]);
                }
                tp_commit_both(&tmem_full[acc]);
            }
        }
    } else {
        // ---- epilogue warps of THIS CTA: its 128 rows x 256 columns (same split of lanes / columns as the single-CTA kernel)
        const int q = warp & 3;
        const int c_lo = warp < 6 ? 0 : TP_BN / 2, c_hi = c_lo + TP_BN / 2;
        uint32_t ti = 0;
        float* sbias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full) + 256);
        bf16* stage = reinterpret_cast<bf16*>(reinterpret_cast<uint8_t*>(full) + 256 + 2048) + (warp - 2) * (32 * 40);   // per-warp [32][40] bf16
        for (int tile = pair; tile < n_tiles; tile += n_pairs, ++ti) {
            const int m0 = (tile % tiles_m) * 2 * TC_BM + (int)rank * TC_BM, n0 = (tile / tiles_m) * TP_BN;
            const uint32_t acc = ti & 1;
            epilogue_tile<EPI>(p, tmem_base + acc * TP_BN, q, lane, c_lo, c_hi, m0, n0, TP_BN, sbias + (ti & 1) * 256, stage, &tmem_full[acc], (ti >> 1) & 1, [&]() {
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    const uint32_t bar = tp_map_to_cta(&tmem_empty[acc], 0);
                    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
                }
            });
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    // the peer may still be signalling into / reading from this CTA's shared memory
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
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

int make_map(CUtensorMap* map, const bf16* base, int rows, int cols, int ld, int box_rows) {
    EncodeTiledFn enc = get_encode();
    PADT_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PADT_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (rc " + std::to_string((int)r) + ")");
    return 0;
}

int sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    }
    return n;
}

// fp32 partial-tile workspace + per-tile arrival counters of the split-K round, one per (device, stream): launches on one stream are
// ordered, so they can share it; the counters return to zero at the end of every launch
int split_workspace(cudaStream_t st, float** ws, unsigned int** counters) {
    static std::mutex mu;
    static std::map<std::pair<int, cudaStream_t>, void*> cache;
    int dev = 0;
    PADT_CHECK_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    void*& buf = cache[std::make_pair(dev, st)];
    const size_t ws_bytes = (size_t)sm_count() * TC_BM * 256 * sizeof(float), cnt_bytes = 4096;
    if (!buf) {
        PADT_CHECK_CUDA(cudaMalloc(&buf, ws_bytes + cnt_bytes));
        PADT_CHECK_CUDA(cudaMemsetAsync(static_cast<uint8_t*>(buf) + ws_bytes, 0, cnt_bytes, st));
    }
    *ws = static_cast<float*>(buf);
    *counters = reinterpret_cast<unsigned int*>(static_cast<uint8_t*>(buf) + ws_bytes);
    return 0;
}

template <int BN, int EPI>
int launch_bn(const GemmParams& p, cudaStream_t st) {
    static bool configured_dev[64] = {false};
    bool& configured = configured_dev[padt_dev_slot()];
    if (!configured) {
        PADT_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem<BN>::TOTAL));
        PADT_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_persistent_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem<BN>::TOTAL));
        configured = true;
    }
    CUtensorMap ma, mw;
    RC_RET(make_map(&ma, p.A, p.M, p.K, p.lda, TC_BM));
    RC_RET(make_map(&mw, p.W, p.N, p.K, p.ldw, BN));
    const int tiles_n = (p.N + BN - 1) / BN, tiles_m = (p.M + TC_BM - 1) / TC_BM;
    if (g_padt_gemm_tc == 2) {      // one tile per CTA (first tcgen05 version, kept for A/B measurements)
        dim3 grid(tiles_n, tiles_m);
        gemm_tc_kernel<BN, EPI><<<grid, TC_THREADS, TcSmem<BN>::TOTAL, st>>>(ma, mw, p);
    } else {
        const int n_tiles = tiles_n * tiles_m, sms = sm_count();
        const int grid = n_tiles < sms ? n_tiles : sms;
        // tiles past the last full round: when they would occupy at most half of the SMs, split their K range so that the round fills
        // the machine (M = 2350, N = 2048: 152 tiles on 148 SMs is two rounds for four tiles' worth of work)
        TcSplit sp = {n_tiles, 0, 1, nullptr, nullptr};
        static const int split_on = getenv("PADT_GEMM_SPLITK") ? atoi(getenv("PADT_GEMM_SPLITK")) : 1;
        const int rest = n_tiles % sms, nk = (p.K + TC_BK - 1) / TC_BK;
        if (split_on && n_tiles > sms && rest > 0 && rest * 2 <= sms && nk >= 48) {
            int parts = sms / rest;
            if (parts > 8) parts = 8;
            if (parts > nk / 16) parts = nk / 16;    // a part is worth its reduction (fp32 round trip through L2, a few us) from ~16 k-blocks up
            if (parts >= 2) {
                float* ws = nullptr;
                unsigned int* counters = nullptr;
                RC_RET(split_workspace(st, &ws, &counters));
                sp.dp_tiles = n_tiles - rest; sp.n_units = rest * parts; sp.parts = parts; sp.ws = ws; sp.counters = counters;
            }
        }
        gemm_tc_persistent_kernel<BN, EPI><<<grid, TC_THREADS_P, TcSmem<BN>::TOTAL, st>>>(ma, mw, p, tiles_m, n_tiles, sp);
    }
    PADT_LAUNCH_CHECK();
    return 0;
}

// Tile width: 256 columns whenever N allows.  Measured (profiles/r01b/call38_gemm_bn_heuristic.log): switching to 128-column tiles
// where they shorten the tile-round count (e.g. M = 2350, N = 2048: 152 tiles of 256 columns on 148 SMs = 2 rounds vs 304 tiles of
// 128 = 1.5 rounds) is SLOWER (down_proj 137 -> 159 us): the narrower tile moves 1.33x the operand bytes per flop through L2.
// PADT_GEMM_BN=128 forces the narrow tile for measurements.
template <int EPI>
int launch_pair(const GemmParams& p, cudaStream_t st) {
    static bool configured_dev[64] = {false};
    bool& configured = configured_dev[padt_dev_slot()];
    if (!configured) {
        PADT_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_pair_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, TP_SMEM));
        configured = true;
    }
    CUtensorMap ma, mw;
    RC_RET(make_map(&ma, p.A, p.M, p.K, p.lda, TC_BM));
    RC_RET(make_map(&mw, p.W, p.N, p.K, p.ldw, TP_BN / 2));
    const int tiles_n = (p.N + TP_BN - 1) / TP_BN, tiles_m = (p.M + 2 * TC_BM - 1) / (2 * TC_BM);
    const int n_tiles = tiles_n * tiles_m, pairs = sm_count() / 2;
    const int grid = 2 * (n_tiles < pairs ? n_tiles : pairs);
    gemm_tc_pair_kernel<EPI><<<grid, TC_THREADS_P, TP_SMEM, st>>>(ma, mw, p, tiles_m, n_tiles);
    PADT_LAUNCH_CHECK();
    return 0;
}

// Tile width: 256 columns whenever N allows.  Measured (profiles/r01b/call38_gemm_bn_heuristic.log): switching to 128-column tiles
// where they shorten the tile-round count (e.g. M = 2350, N = 2048: 152 tiles of 256 columns on 148 SMs = 2 rounds vs 304 tiles of
// 128 = 1.5 rounds) is SLOWER (down_proj 137 -> 159 us): the narrower tile moves 1.33x the operand bytes per flop through L2.
// PADT_GEMM_BN=128 forces the narrow tile for measurements.
// CTA pairs (256 x 256 tiles, cta_group::2) when their round quantisation does not cost more than the pair's faster mainloop gains:
// PADT_GEMM_PAIR = 0 never, 2 whenever the shape allows, default 1 = by the estimate below.
template <int EPI>
int launch_epi(const GemmParams& p, cudaStream_t st) {
    static const int forced = getenv("PADT_GEMM_BN") ? atoi(getenv("PADT_GEMM_BN")) : 0;
    static const int pair_mode = getenv("PADT_GEMM_PAIR") ? atoi(getenv("PADT_GEMM_PAIR")) : 1;
    if (p.N < 256 || forced == 128) return launch_bn<128, EPI>(p, st);
    if (pair_mode && p.M >= 512 && forced != 256) {
        const int sms = sm_count(), pairs = sms / 2;
        const long long t1 = (long long)((p.M + TC_BM - 1) / TC_BM) * ((p.N + 255) / 256);
        const long long t2 = (long long)((p.M + 2 * TC_BM - 1) / (2 * TC_BM)) * ((p.N + 255) / 256);
        const int nk = (p.K + TC_BK - 1) / TC_BK;
        // rounds in units of one 128 x 256 tile time; the single-CTA kernel K-splits a short last round when K is long (launch_bn)
        double r1 = (double)((t1 + sms - 1) / sms);
        if (nk >= 48 && t1 > sms && (t1 % sms) * 2 <= sms && t1 % sms) r1 = (double)(t1 / sms) + 0.6;
        const double r2 = (double)((t2 + pairs - 1) / pairs) * 0.80;      // pair mainloop: ~0.8 of the single-CTA tile time (measured)
        // the SwiGLU epilogue is math-bound and gains nothing from the pair (measured: 159.7 vs 172 us at the LLM gate/up shape)
        if (pair_mode == 2 || (r2 < r1 && EPI != EPI_SWIGLU)) return launch_pair<EPI>(p, st);
    }
    return launch_bn<256, EPI>(p, st);
}

}  // namespace

// Shapes the tcgen05 path takes; everything else (small M / N / K, odd strides) stays on the mma.sync kernel.
bool gemm_tc_eligible(const GemmParams& p) {
    return p.M >= 256 && p.N >= 128 && p.K >= 64 && (p.lda % 8) == 0 && (p.ldw % 8) == 0 && (((uintptr_t)p.A) & 15) == 0 &&
           (((uintptr_t)p.W) & 15) == 0 && get_encode() != nullptr;
}

int launch_gemm_tc(int epi, const GemmParams& p, cudaStream_t st) {
    switch (epi) {
        case EPI_NONE: return launch_epi<EPI_NONE>(p, st);
        case EPI_GELU: return launch_epi<EPI_GELU>(p, st);
        case EPI_RESID: return launch_epi<EPI_RESID>(p, st);
        case EPI_SWIGLU: return launch_epi<EPI_SWIGLU>(p, st);
    }
    padt_set_error("launch_gemm_tc: unknown epilogue");
    return -1;
}
