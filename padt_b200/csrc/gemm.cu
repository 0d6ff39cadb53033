// Dense projection  C[M,N] = A[M,K] * W[N,K]^T (+bias) with fused epilogues   (bf16 in, fp32 accumulate)
//
// Replaces every nn.Linear call of the reference's tensor-core-bound stages (cuBLAS via PyTorch in the
// reference): ViT patch-embed / qkv / proj / MLP / merger (HF:91-146,207-322),
// LLM prefill projections (HF:611-624,672-759), PaDT decoder projections (padt_decoder.py:11-60,82-86,142-184).
// Rounding points are the HF ones: Linear output rounded to bf16 once (acc + bias in fp32), activation on the
// rounded value, residual add rounded again.
//
// v1 kernel: mma.sync m16n8k16 tiles fed by a 4-stage cp.async pipeline, 128x128x32 CTA tile, XOR-swizzled smem.
// (The tcgen05/TMA replacement for the large-M shapes is gemm_tc.cu.)
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int BM = 128, BN = 128, BK = 32, STAGES = 4, THREADS = 256;
constexpr int SMEM_BYTES = STAGES * (BM + BN) * BK * 2;

__device__ __forceinline__ int swz(int row, int ch) { return ch ^ ((row >> 1) & 3); }

template <int EPI>
__global__ void __launch_bounds__(THREADS) gemm_bf16_kernel(GemmParams p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    bf16* sA = reinterpret_cast<bf16*>(smem_raw);
    bf16* sB = sA + STAGES * BM * BK;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 2, wn = warp & 3;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int nk = (p.K + BK - 1) / BK;

    auto load_stage = [&](int s, int kt) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            int id = tid + i * THREADS;
            int row = id >> 2, ch = id & 3;
            int gk = kt * BK + ch * 8;
            int kb = (p.K - gk) * 2;
            kb = kb < 0 ? 0 : (kb > 16 ? 16 : kb);
            {
                int gr = m0 + row;
                int bytes = gr < p.M ? kb : 0;
                const bf16* src = p.A + (size_t)(gr < p.M ? gr : 0) * p.lda + (bytes > 0 ? gk : 0);
                cp_async16(sA + (s * BM + row) * BK + swz(row, ch) * 8, src, bytes);
            }
            {
                int gr = n0 + row;
                int bytes = gr < p.N ? kb : 0;
                const bf16* src = p.W + (size_t)(gr < p.N ? gr : 0) * p.ldw + (bytes > 0 ? gk : 0);
                cp_async16(sB + (s * BN + row) * BK + swz(row, ch) * 8, src, bytes);
            }
        }
    };

    float acc[4][4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) load_stage(s, s);
        cp_async_commit();
    }

    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int nx = kt + STAGES - 1;
            if (nx < nk) load_stage(nx % STAGES, nx);
            cp_async_commit();
        }
        const bf16* a_s = sA + (kt % STAGES) * BM * BK;
        const bf16* b_s = sB + (kt % STAGES) * BN * BK;
#pragma unroll
        for (int kk = 0; kk < BK / 16; ++kk) {
            uint32_t af[4][4], bfr[2][4];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) {
                int row = wm * 64 + mi * 16 + (lane & 15);
                int ch = kk * 2 + (lane >> 4);
                ldmatrix_x4(af[mi][0], af[mi][1], af[mi][2], af[mi][3], a_s + row * BK + swz(row, ch) * 8);
            }
#pragma unroll
            for (int nj = 0; nj < 2; ++nj) {
                int row = wn * 32 + nj * 16 + (lane & 7) + ((lane >> 4) << 3);
                int ch = kk * 2 + ((lane >> 3) & 1);
                ldmatrix_x4(bfr[nj][0], bfr[nj][1], bfr[nj][2], bfr[nj][3], b_s + row * BK + swz(row, ch) * 8);
            }
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni)
                    mma_bf16_16816(acc[mi][ni], af[mi], bfr[ni >> 1][(ni & 1) * 2], bfr[ni >> 1][(ni & 1) * 2 + 1]);
        }
    }
    cp_async_wait<0>();

    // ---- epilogue
    const int n_out = (EPI == EPI_SWIGLU) ? p.N / 2 : p.N;
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            int row = m0 + wm * 64 + mi * 16 + (lane >> 2) + half * 8;
            if (row >= p.M) continue;
            int orow = p.out_row_map ? p.out_row_map[row] : row;
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                int col = n0 + wn * 32 + ni * 8 + (lane & 3) * 2;
                if (col >= p.N) continue;
                float v0 = acc[mi][ni][half * 2 + 0], v1 = acc[mi][ni][half * 2 + 1];
                bool has1 = (col + 1) < p.N;
                if (p.bias) {
                    v0 += bf2f(p.bias[col]);
                    if (has1) v1 += bf2f(p.bias[col + 1]);
                }
                v0 = rbf(v0);
                v1 = rbf(v1);
                if (EPI == EPI_SWIGLU) {
                    // interleaved rows: even = gate_j, odd = up_j  ->  bf16(bf16(silu(g)) * u)   (HF:88 / HF:623)
                    float r = rbf(rbf(silu_f(v0)) * v1);
                    p.C[(size_t)orow * p.ldc + (col >> 1)] = f2bf(r);
                    continue;
                }
                if (EPI == EPI_GELU) {
                    v0 = rbf(gelu_erf_f(v0));
                    v1 = rbf(gelu_erf_f(v1));
                }
                if (EPI == EPI_RESID) {
                    const bf16* r = p.resid + (size_t)orow * p.ldr + col;
                    v0 = rbf(bf2f(r[0]) + v0);
                    if (has1) v1 = rbf(bf2f(r[1]) + v1);
                }
                bf16* dst = p.C + (size_t)orow * p.ldc + col;
                if (has1 && ((p.ldc & 1) == 0)) {
                    *reinterpret_cast<uint32_t*>(dst) = f2_to_bf2x(v0, v1);
                } else {
                    dst[0] = f2bf(v0);
                    if (has1) dst[1] = f2bf(v1);
                }
            }
        }
    }
    (void)n_out;
}

// ---------------------------------------------------------------------------------------------------------------------------------
// Skinny path: M <= 16 activation rows per row block (the PaDT decoder's query side: 3 + n_vrt rows per object, its heads: one row per
// object).  The tiled kernel above launches N/64 = 20 CTAs that each walk K serially (~25 us for a 1280 x 1280 Linear whose 3.3 MB of
// weights stream in ~1 us): 39 such launches were 1.1 ms of a 2.5 ms vl_decode (profiles/r02/call29).  Here every warp owns two weight
// rows, the lanes stride over K in 16-byte chunks against the activation rows held in shared memory (same scheme as the decode GEMV),
// N/16 CTAs per row block.  Same epilogue arithmetic and rounding points as gemm_bf16_kernel.
constexpr int SK_THREADS = 256, SK_WARPS = SK_THREADS / 32, SK_MB = 16, SK_UNR = 4;

template <int EPI>
__global__ void __launch_bounds__(SK_THREADS) gemm_skinny_kernel(GemmParams p) {
    extern __shared__ __align__(16) uint8_t sk_raw[];
    bf16* xs = reinterpret_cast<bf16*>(sk_raw);                 // [SK_MB][Kp], Kp = K rounded up to 8, tail zeroed
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Kp = (p.K + 7) & ~7, nch = Kp >> 3;
    const int m0 = blockIdx.y * SK_MB;
    const int mrows = p.M - m0 < SK_MB ? p.M - m0 : SK_MB;
    for (int i = tid; i < SK_MB * nch; i += SK_THREADS) {
        const int r = i / nch, c = i % nch;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (r < mrows) {
            v = *reinterpret_cast<const uint4*>(p.A + (size_t)(m0 + r) * p.lda + c * 8);
            if (c * 8 + 8 > p.K) {                               // K tail: whatever sits in the pad of A must not count
                bf16* e = reinterpret_cast<bf16*>(&v);
                for (int k = 0; k < 8; ++k)
                    if (c * 8 + k >= p.K) e[k] = f2bf(0.f);
            }
        }
        *reinterpret_cast<uint4*>(xs + (size_t)r * Kp + c * 8) = v;
    }
    __syncthreads();
    const int n_pairs = (p.N + 1) >> 1;
    for (int pair = blockIdx.x * SK_WARPS + warp; pair < n_pairs; pair += gridDim.x * SK_WARPS) {
        const int n0 = pair * 2, n1 = n0 + 1;
        const bool has1 = n1 < p.N;
        const bf16* w0 = p.W + (size_t)n0 * p.ldw;
        const bf16* w1 = p.W + (size_t)(has1 ? n1 : n0) * p.ldw;
        float a0[SK_MB], a1[SK_MB];
#pragma unroll
        for (int r = 0; r < SK_MB; ++r) { a0[r] = 0.f; a1[r] = 0.f; }
        // four 16-byte chunks of both rows in flight per lane before the first FMA (K = 1280: 5 dependent round trips to HBM became 2);
        // the accumulation order per lane is unchanged (c = lane, lane + 32, ...)
        for (int c0 = lane; c0 < nch; c0 += 32 * SK_UNR) {
            uint4 wav[SK_UNR], wbv[SK_UNR];
#pragma unroll
            for (int u = 0; u < SK_UNR; ++u) {
                const int c = c0 + u * 32;
                wav[u] = c < nch ? ldg_nc_v4(w0 + (size_t)c * 8) : make_uint4(0, 0, 0, 0);
                wbv[u] = c < nch ? ldg_nc_v4(w1 + (size_t)c * 8) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < SK_UNR; ++u) {
                const int c = c0 + u * 32;
                if (c >= nch) break;
                const uint4 wa = wav[u], wb = wbv[u];
                const float2 wa0 = bf2x_to_f2(wa.x), wa1 = bf2x_to_f2(wa.y), wa2 = bf2x_to_f2(wa.z), wa3 = bf2x_to_f2(wa.w);
                const float2 wb0 = bf2x_to_f2(wb.x), wb1 = bf2x_to_f2(wb.y), wb2 = bf2x_to_f2(wb.z), wb3 = bf2x_to_f2(wb.w);
#pragma unroll
                for (int r = 0; r < SK_MB; ++r) {
                    if (r >= mrows) break;
                    const uint4 xv = *reinterpret_cast<const uint4*>(xs + (size_t)r * Kp + c * 8);
                    const float2 x0 = bf2x_to_f2(xv.x), x1 = bf2x_to_f2(xv.y), x2 = bf2x_to_f2(xv.z), x3 = bf2x_to_f2(xv.w);
                    float s0 = a0[r], s1 = a1[r];
                    s0 = fmaf(wa0.x, x0.x, s0); s0 = fmaf(wa0.y, x0.y, s0); s0 = fmaf(wa1.x, x1.x, s0); s0 = fmaf(wa1.y, x1.y, s0);
                    s0 = fmaf(wa2.x, x2.x, s0); s0 = fmaf(wa2.y, x2.y, s0); s0 = fmaf(wa3.x, x3.x, s0); s0 = fmaf(wa3.y, x3.y, s0);
                    s1 = fmaf(wb0.x, x0.x, s1); s1 = fmaf(wb0.y, x0.y, s1); s1 = fmaf(wb1.x, x1.x, s1); s1 = fmaf(wb1.y, x1.y, s1);
                    s1 = fmaf(wb2.x, x2.x, s1); s1 = fmaf(wb2.y, x2.y, s1); s1 = fmaf(wb3.x, x3.x, s1); s1 = fmaf(wb3.y, x3.y, s1);
                    a0[r] = s0; a1[r] = s1;
                }
            }
        }
#pragma unroll
        for (int r = 0; r < SK_MB; ++r) {
            if (r >= mrows) break;
            float v0 = warp_sum(a0[r]), v1 = warp_sum(a1[r]);
            if (lane != 0) continue;
            const int row = m0 + r;
            if (p.bias) { v0 += bf2f(p.bias[n0]); if (has1) v1 += bf2f(p.bias[n1]); }
            v0 = rbf(v0); v1 = rbf(v1);
            if (EPI == EPI_GELU) { v0 = rbf(gelu_erf_f(v0)); v1 = rbf(gelu_erf_f(v1)); }
            if (EPI == EPI_RESID) {
                const bf16* rs = p.resid + (size_t)row * p.ldr + n0;
                v0 = rbf(bf2f(rs[0]) + v0);
                if (has1) v1 = rbf(bf2f(rs[1]) + v1);
            }
            bf16* dst = p.C + (size_t)row * p.ldc + n0;
            dst[0] = f2bf(v0);
            if (has1) dst[1] = f2bf(v1);
        }
    }
}

template <int EPI>
int launch_skinny(const GemmParams& p, cudaStream_t st) {
    const int Kp = (p.K + 7) & ~7;
    const size_t smem = (size_t)SK_MB * Kp * sizeof(bf16);
    static size_t configured_dev[64] = {0};
    size_t& configured = configured_dev[padt_dev_slot()];
    if (smem > configured) {
        PADT_CHECK_CUDA(cudaFuncSetAttribute(gemm_skinny_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem > 48 * 1024 ? smem : 48 * 1024)));
        configured = smem > 48 * 1024 ? smem : 48 * 1024;
    }
    int gx = cdiv((p.N + 1) / 2, SK_WARPS);
    gx = gx > 148 * 4 ? 148 * 4 : gx;
    dim3 grid(gx, cdiv(p.M, SK_MB));
    gemm_skinny_kernel<EPI><<<grid, SK_THREADS, smem, st>>>(p);
    PADT_LAUNCH_CHECK();
    return 0;
}

template <int EPI>
int launch_t(const GemmParams& p, cudaStream_t st) {
    static bool configured_dev[64] = {false};
    bool& configured = configured_dev[padt_dev_slot()];
    if (!configured) {
        PADT_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        configured = true;
    }
    dim3 grid(cdiv(p.N, BN), cdiv(p.M, BM));
    gemm_bf16_kernel<EPI><<<grid, THREADS, SMEM_BYTES, st>>>(p);
    PADT_LAUNCH_CHECK();
    return 0;
}

}  // namespace

int g_padt_gemm_tc = 1;

int launch_gemm(int epi, const GemmParams& p_in, cudaStream_t st) {
    if (p_in.M == 0 || p_in.N == 0) return 0;
    GemmParams p = p_in;
    static const int noepi = getenv("PADT_GEMM_NOEPI") ? atoi(getenv("PADT_GEMM_NOEPI")) : 0;
    p.dbg_skip_epilogue = noepi;
    if (const char* e = getenv("PADT_GEMM_TC")) g_padt_gemm_tc = atoi(e);   // A/B switch: 0 mma.sync, 1 persistent tcgen05, 2 one-tile tcgen05
    if (g_padt_gemm_tc && gemm_tc_eligible(p) && (epi != EPI_SWIGLU || (p.N % 2) == 0) && (epi != EPI_RESID || p.resid != nullptr))
        return launch_gemm_tc(epi, p, st);
    PADT_REQUIRE(p.K > 0, "K");
    PADT_REQUIRE((p.lda % 8) == 0 && (p.ldw % 8) == 0, "lda/ldw must be multiples of 8 elements (16-byte rows)");
    PADT_REQUIRE((((uintptr_t)p.A) & 15) == 0 && (((uintptr_t)p.W) & 15) == 0, "A/W must be 16-byte aligned");
    PADT_REQUIRE(((uintptr_t)p.C & 3) == 0, "C must be 4-byte aligned");
    // few activation rows (decoder query side / heads): the skinny path; in-place residual (C == resid) is fine, every element is read
    // and written by the same lane
    const bool skinny_ok = p.M <= 64 && !p.out_row_map && epi != EPI_SWIGLU && (size_t)SK_MB * ((p.K + 7) & ~7) * 2 <= 200 * 1024 &&
                           !(getenv("PADT_GEMM_SKINNY") && atoi(getenv("PADT_GEMM_SKINNY")) == 0);
    if (skinny_ok) {
        switch (epi) {
            case EPI_NONE: return launch_skinny<EPI_NONE>(p, st);
            case EPI_GELU: return launch_skinny<EPI_GELU>(p, st);
            case EPI_RESID:
                PADT_REQUIRE(p.resid != nullptr, "resid");
                return launch_skinny<EPI_RESID>(p, st);
        }
    }
    switch (epi) {
        case EPI_NONE: return launch_t<EPI_NONE>(p, st);
        case EPI_GELU: return launch_t<EPI_GELU>(p, st);
        case EPI_RESID:
            PADT_REQUIRE(p.resid != nullptr, "resid");
            return launch_t<EPI_RESID>(p, st);
        case EPI_SWIGLU:
            PADT_REQUIRE((p.N % 2) == 0, "interleaved gate/up needs even N");
            return launch_t<EPI_SWIGLU>(p, st);
    }
    padt_set_error("launch_gemm: unknown epilogue");
    return -1;
}
