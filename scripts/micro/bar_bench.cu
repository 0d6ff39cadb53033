// Micro-benchmark: latency of one grid-wide barrier on B200 for a few implementations (148 CTAs x 544 threads, cooperative launch).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/bar_bench scripts/micro/bar_bench.cu && /tmp/bar_bench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }

template <int V>
__global__ void __launch_bounds__(544, 1) bar_kernel(unsigned* counter, unsigned* flags, float* data, int iters, long long* out, const float* stream, size_t stream_elems, int with_traffic) {
    const int tid = threadIdx.x, G = gridDim.x, cta = blockIdx.x;
    unsigned target = 0;
    float acc = 0.f;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        // a little "epilogue": every CTA publishes a few values that others read after the barrier
        if (tid < 16) data[cta * 16 + tid] = (float)(it + tid);
        if (with_traffic && tid >= 512) {   // one warp keeps loads in flight (stand-in for the TMA weight stream)
            size_t base = ((size_t)cta * 977 + (size_t)it * 131071) * 1024 % (stream_elems - 32 * 64);
            for (int k = 0; k < 16; ++k) acc += __ldcs(stream + base + k * 2048 + (tid - 512) * 4);
        }
        __syncthreads();
        if (V == 0) {
            target += G;
            if (tid == 0) {
                asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
                while (ld_acquire(counter) < target) {}
            }
        } else if (V == 1) {           // relaxed polls, one acquire fence at the end
            target += G;
            if (tid == 0) {
                asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
                while (ld_relaxed(counter) < target) {}
                asm volatile("fence.acq_rel.gpu;" ::: "memory");
            }
        } else if (V == 2) {           // per-CTA flags, no atomics: warp 0 polls all flags
            target += 1;
            if (tid == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + cta * 32), "r"(target) : "memory");
            if (tid < 32) {
                for (int c = tid; c < G; c += 32) while (ld_relaxed(flags + c * 32) < target) {}
                asm volatile("fence.acq_rel.gpu;" ::: "memory");
            }
        } else if (V == 3) {           // explicit fence + relaxed red, relaxed polls
            target += G;
            if (tid == 0) {
                asm volatile("fence.acq_rel.gpu;" ::: "memory");
                asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
                while (ld_relaxed(counter) < target) {}
                asm volatile("fence.acq_rel.gpu;" ::: "memory");
            }
        } else if (V == 4) {           // atomics spread over 4 counters (one per CTA quarter), poll all four
            target += 1;
            if (tid == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter + (cta & 3) * 32) : "memory");
            if (tid < 4) {
                unsigned want = target * ((G + 3 - tid) / 4);
                while (ld_relaxed(counter + tid * 32) < want) {}
                asm volatile("fence.acq_rel.gpu;" ::: "memory");
            }
        }
        __syncthreads();
        // consume something another CTA wrote (keeps the compiler and the memory model honest)
        if (tid < 16) acc += __ldcg(data + ((cta + 1 + it) % G) * 16 + tid);
    }
    long long t1 = clock64();
    if (tid == 0) out[cta] = t1 - t0;
    if (acc == 12345.678f) out[0] = 0;
}

int main() {
    int dev = 0; cudaSetDevice(dev);
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, dev);
    int G = prop.multiProcessorCount, iters = 2000;
    unsigned *counter, *flags; float* data; long long* out; float* stream;
    size_t stream_elems = (size_t)1 << 28;   // 1 GiB of floats
    cudaMalloc(&counter, 4096); cudaMalloc(&flags, G * 128); cudaMalloc(&data, G * 64); cudaMalloc(&out, G * 8); cudaMalloc(&stream, stream_elems * 4);
    cudaMemset(stream, 0, stream_elems * 4);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    for (int traffic = 0; traffic < 2; ++traffic)
        for (int v = 0; v < 5; ++v) {
            cudaMemset(counter, 0, 4096); cudaMemset(flags, 0, G * 128);
            void* args[] = {&counter, &flags, &data, &iters, &out, &stream, &stream_elems, &traffic};
            const void* fn = v == 0 ? (const void*)bar_kernel<0> : v == 1 ? (const void*)bar_kernel<1> : v == 2 ? (const void*)bar_kernel<2> : v == 3 ? (const void*)bar_kernel<3> : (const void*)bar_kernel<4>;
            cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3(G), dim3(544), args, 0, 0);
            if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return 1; }
            e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("run failed: %s\n", cudaGetErrorString(e)); return 1; }
            long long h[256]; cudaMemcpy(h, out, G * 8, cudaMemcpyDeviceToHost);
            printf("{\"variant\": %d, \"traffic\": %d, \"cycles_per_barrier\": %.1f, \"us_per_barrier_at_%dMHz\": %.3f}\n", v, traffic, (double)h[0] / iters, clk / 1000,
                   (double)h[0] / iters / (clk / 1000.0));
        }
    return 0;
}
