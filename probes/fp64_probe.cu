// FP64 pipe probe for B200 (sm_100a): DFMA peak, DMMA (mma.sync m8n8k4 f64) peak,
// mixed issue, and SM clock under load.  Establishes the FP64 roofline denominator
// that MEASURED_PEAKS.json does not carry.  Build: see probes/Makefile.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void __launch_bounds__(512) k_dmma(double *out, int iters, double a0, double b0) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; i++) { c[i][0] = 0.0; c[i][1] = 0.0; }
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(512) k_dfma(double *out, int iters, double a0, double b0) {
    double c[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = i * 1e-3;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Mixed: per loop iteration NM DMMAs and NF DFMAs (independent chains) from the same warp.
template <int NM, int NF>
__global__ void __launch_bounds__(512) k_mixed(double *out, int iters, double a0, double b0) {
    double c[NM > 0 ? NM : 1][2];
    double f[NF > 0 ? NF : 1];
#pragma unroll
    for (int i = 0; i < NM; i++) { c[i][0] = 0.0; c[i][1] = 0.0; }
#pragma unroll
    for (int i = 0; i < NF; i++) f[i] = i * 1e-3;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NM; i++) dmma884(c[i][0], c[i][1], a, b);
#pragma unroll
        for (int i = 0; i < NF; i++) f[i] = fma(f[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NM; i++) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < NF; i++) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Warp-specialised mix: even warps DMMA, odd warps DFMA (same SMSP hosts both kinds when 8+ warps)
__global__ void __launch_bounds__(512) k_split(double *out, int iters, double a0, double b0, int fma_warps_mask) {
    int warp = threadIdx.x >> 5;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    double s = 0;
    if ((fma_warps_mask >> (warp & 7)) & 1) {
        double f[12];
#pragma unroll
        for (int i = 0; i < 12; i++) f[i] = i * 1e-3;
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 12; i++) f[i] = fma(f[i], a, b);
        }
#pragma unroll
        for (int i = 0; i < 12; i++) s += f[i];
    } else {
        double c[12][2];
#pragma unroll
        for (int i = 0; i < 12; i++) { c[i][0] = 0; c[i][1] = 0; }
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 12; i++) dmma884(c[i][0], c[i][1], a, b);
        }
#pragma unroll
        for (int i = 0; i < 12; i++) s += c[i][0] + c[i][1];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_clock(long long *out, int spin) {
    long long t0 = clock64();
    unsigned long long g0, g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    double x = 1.0;
    for (int i = 0; i < spin; i++) x = fma(x, 1.0000001, 1e-9);
    long long t1 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = (long long)(g1 - g0); out[2] = (long long)x; }
}

template <typename F>
static float timeit(F f, int reps = 5) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device: %s sm_%d%d SMs=%d smem/block optin=%zu regs/SM=%d L2=%d MB clock=%d kHz mem=%zu MB\n",
           p.name, p.major, p.minor, p.multiProcessorCount, p.sharedMemPerBlockOptin, p.regsPerMultiprocessor,
           p.l2CacheSize >> 20, p.clockRate, p.totalGlobalMem >> 20);
    int nsm = p.multiProcessorCount;
    double *out; CK(cudaMalloc(&out, sizeof(double) * nsm * 8 * 512));
    long long *clk; CK(cudaMalloc(&clk, 64));
    const int iters = 20000;

    for (int threads : {128, 256, 512}) {
        for (int cta_per_sm : {1, 2}) {
            if (threads * cta_per_sm > 1024 && cta_per_sm > 1 && threads == 512) {}
            int grid = nsm * cta_per_sm;
            float ms = timeit([&] { k_dmma<12><<<grid, threads>>>(out, iters, 1.0, 1e-9); });
            double flops = (double)grid * (threads / 32) * iters * 12.0 * 512.0;
            printf("DMMA m8n8k4 NACC=12 threads=%d cta/sm=%d : %.3f ms  %.2f TFLOP/s\n", threads, cta_per_sm, ms, flops / ms / 1e9);
            ms = timeit([&] { k_dfma<12><<<grid, threads>>>(out, iters, 1.0000001, 1e-9); });
            flops = (double)grid * threads * iters * 12.0 * 2.0;
            printf("DFMA        NACC=12 threads=%d cta/sm=%d : %.3f ms  %.2f TFLOP/s\n", threads, cta_per_sm, ms, flops / ms / 1e9);
        }
    }
    {   // ILP sensitivity of DMMA at 16 warps/SM
        int grid = nsm, threads = 512;
        float ms;
        ms = timeit([&] { k_dmma<1><<<grid, threads>>>(out, iters, 1.0, 1e-9); });
        printf("DMMA NACC=1  512thr: %.2f TFLOP/s (dep-chain latency = %.1f ns/op)\n", (double)grid * 16 * iters * 1 * 512.0 / ms / 1e9, ms * 1e6 / iters);
        ms = timeit([&] { k_dmma<2><<<grid, threads>>>(out, iters, 1.0, 1e-9); });
        printf("DMMA NACC=2  512thr: %.2f TFLOP/s\n", (double)grid * 16 * iters * 2 * 512.0 / ms / 1e9);
        ms = timeit([&] { k_dmma<4><<<grid, threads>>>(out, iters, 1.0, 1e-9); });
        printf("DMMA NACC=4  512thr: %.2f TFLOP/s\n", (double)grid * 16 * iters * 4 * 512.0 / ms / 1e9);
        ms = timeit([&] { k_dmma<24><<<grid, threads>>>(out, iters, 1.0, 1e-9); });
        printf("DMMA NACC=24 512thr: %.2f TFLOP/s\n", (double)grid * 16 * iters * 24 * 512.0 / ms / 1e9);
        ms = timeit([&] { k_dmma<1><<<grid, 32>>>(out, iters, 1.0, 1e-9); });
        printf("DMMA NACC=1 1 warp/SM: latency %.1f ns per dependent DMMA\n", ms * 1e6 / iters);
        ms = timeit([&] { k_dfma<1><<<grid, 32>>>(out, iters, 1.0000001, 1e-9); });
        printf("DFMA NACC=1 1 warp/SM: latency %.1f ns per dependent DFMA\n", ms * 1e6 / iters);
    }
    {   // mixed in one warp
        int grid = nsm, threads = 512;
        float ms;
        ms = timeit([&] { k_mixed<12, 0><<<grid, threads>>>(out, iters, 1.0, 1e-9); });
        printf("mixed 12 DMMA + 0 DFMA : %.3f ms\n", ms);
        ms = timeit([&] { k_mixed<12, 12><<<grid, threads>>>(out, iters, 1.0, 1e-9); });
        printf("mixed 12 DMMA + 12 DFMA: %.3f ms (DMMA flops %.2f TF, DFMA flops %.2f TF)\n", ms,
               (double)grid * 16 * iters * 12 * 512.0 / ms / 1e9, (double)grid * 512 * iters * 12 * 2.0 / ms / 1e9);
        ms = timeit([&] { k_mixed<12, 24><<<grid, threads>>>(out, iters, 1.0, 1e-9); });
        printf("mixed 12 DMMA + 24 DFMA: %.3f ms (DMMA flops %.2f TF, DFMA flops %.2f TF)\n", ms,
               (double)grid * 16 * iters * 12 * 512.0 / ms / 1e9, (double)grid * 512 * iters * 24 * 2.0 / ms / 1e9);
        ms = timeit([&] { k_mixed<0, 24><<<grid, threads>>>(out, iters, 1.0, 1e-9); });
        printf("mixed 0 DMMA + 24 DFMA : %.3f ms\n", ms);
        for (int mask : {0x00, 0xAA, 0xF0, 0xFF}) {
            ms = timeit([&] { k_split<<<grid, threads>>>(out, iters, 1.0, 1e-9, mask); });
            printf("warp-split fma_mask=0x%02x: %.3f ms\n", mask, ms);
        }
    }
    {   // SM clock under (light) load and under FP64 load
        k_clock<<<1, 32>>>(clk, 2000000); CK(cudaDeviceSynchronize());
        long long h[3]; CK(cudaMemcpy(h, clk, 24, cudaMemcpyDeviceToHost));
        printf("idle-ish clock: %.1f MHz\n", (double)h[0] / (double)h[1] * 1e3);
        cudaStream_t s2; cudaStreamCreate(&s2);
        k_dmma<12><<<nsm - 1, 512>>>(out, 400000, 1.0, 1e-9);
        k_clock<<<1, 32, 0, s2>>>(clk, 20000000); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, clk, 24, cudaMemcpyDeviceToHost));
        printf("clock under DMMA load: %.1f MHz\n", (double)h[0] / (double)h[1] * 1e3);
        k_dfma<12><<<nsm - 1, 512>>>(out, 400000, 1.0000001, 1e-9);
        k_clock<<<1, 32, 0, s2>>>(clk, 20000000); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, clk, 24, cudaMemcpyDeviceToHost));
        printf("clock under DFMA load: %.1f MHz\n", (double)h[0] / (double)h[1] * 1e3);
    }
    return 0;
}
