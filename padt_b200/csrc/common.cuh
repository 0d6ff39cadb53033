// Shared device/host helpers for the PaDT sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

typedef __nv_bfloat16 bf16;

// ------------------------------------------------------------------------------------------------
// error plumbing: every C-ABI call returns 0 or a negative code and records a message
// ------------------------------------------------------------------------------------------------
void padt_set_error(const std::string& msg);
#define PADT_CHECK_CUDA(expr)                                                                   \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            padt_set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " @" + __FILE__ + ":" + std::to_string(__LINE__)); \
            return -2;                                                                          \
        }                                                                                       \
    } while (0)
#define PADT_REQUIRE(cond, msg)                                                                 \
    do {                                                                                        \
        if (!(cond)) {                                                                          \
            padt_set_error(std::string("requirement failed: ") + #cond + " : " + (msg) + " @" + __FILE__ + ":" + std::to_string(__LINE__)); \
            return -1;                                                                          \
        }                                                                                       \
    } while (0)
extern unsigned long long g_padt_launches;   // kernels launched by this library (graph replays add their node count)
#define PADT_LAUNCH_CHECK()                       \
    do {                                          \
        ++g_padt_launches;                        \
        PADT_CHECK_CUDA(cudaGetLastError());      \
    } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
// index for per-device "already configured" tables: cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-device attribute, and one
// process may own engines on several GPUs
static inline int padt_dev_slot() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) d = 0;
    return d;
}
static inline long long cdivll(long long a, long long b) { return (a + b - 1) / b; }

// ------------------------------------------------------------------------------------------------
// bf16 arithmetic with explicit rounding points (the HF bf16 path rounds after every elementwise op)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float bf2f(bf16 x) { return __bfloat162float(x); }
__device__ __forceinline__ bf16 f2bf(float x) { return __float2bfloat16_rn(x); }
__device__ __forceinline__ float rbf(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }  // round-trip

__device__ __forceinline__ float2 bf2x_to_f2(uint32_t v) {
    // two packed bf16 -> two floats (bf16 is the top half of an fp32)
    float2 r;
    r.x = __uint_as_float(v << 16);
    r.y = __uint_as_float(v & 0xffff0000u);
    return r;
}
__device__ __forceinline__ uint32_t f2_to_bf2x(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + expf(-x)); }
__device__ __forceinline__ float gelu_erf_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// block-wide sum, result broadcast to all threads; `red` must hold >= 33 floats
__device__ __forceinline__ float block_sum(float v, float* red) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    if (w == 0) {
        float t = lane < nw ? red[lane] : 0.f;
        t = warp_sum(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}

// RMSNorm statistics with ONE fixed summation order (lane-strided 8-element chunks, xor-butterfly), shared by the
// row kernel (elementwise.cu) and the decode-step prologues (decode.cu) so both give bit-identical normalised rows.
__device__ __forceinline__ float rms_rstd_warp(const bf16* __restrict__ xr, int dim, float eps, int lane) {
    float ss = 0.f;
    for (int c = lane * 8; c < dim; c += 256) {
        uint4 v = *reinterpret_cast<const uint4*>(xr + c);
        float2 a = bf2x_to_f2(v.x), b = bf2x_to_f2(v.y), cc = bf2x_to_f2(v.z), d = bf2x_to_f2(v.w);
        ss += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y + cc.x * cc.x + cc.y * cc.y + d.x * d.x + d.y * d.y;
    }
    ss = warp_sum(ss);
    return 1.0f / sqrtf(ss / (float)dim + eps);
}
// one 8-element chunk of  w * bf16(x * rstd)   (HF:70-71)
__device__ __forceinline__ uint4 rms_apply8(uint4 v, uint4 wv, float rstd) {
    uint32_t xi[4] = {v.x, v.y, v.z, v.w}, wi[4] = {wv.x, wv.y, wv.z, wv.w}, oi[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float2 xv = bf2x_to_f2(xi[k]), ww = bf2x_to_f2(wi[k]);
        oi[k] = f2_to_bf2x(ww.x * rbf(xv.x * rstd), ww.y * rbf(xv.y * rstd));
    }
    return make_uint4(oi[0], oi[1], oi[2], oi[3]);
}

// ------------------------------------------------------------------------------------------------
// async copy / ldmatrix / mma.sync wrappers (legacy tensor path; the tcgen05 path lives in gemm_tc.cu)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem)), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void ldmatrix_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(smem_u32(p)));
}
// D(16x8,f32) += A(16x16,bf16,row) * B(16x8,bf16,col)
__device__ __forceinline__ void mma_bf16_16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];\n" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
