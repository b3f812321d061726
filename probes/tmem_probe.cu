// TMEM store / load round-trip probe (sm_100a): can the 60-64 accumulator registers of the Gram kernel be parked in
// tensor memory during the J0 phase?  16 warps x 32 lanes x 64 registers = 128 KB per CTA, one CTA per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_probe tmem_probe.cu && ./tmem_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_st64(uint32_t taddr, const uint32_t (&r)[64])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x64.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63, %64};" :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]), "r"(r[32]), "r"(r[33]), "r"(r[34]), "r"(r[35]), "r"(r[36]), "r"(r[37]), "r"(r[38]), "r"(r[39]), "r"(r[40]), "r"(r[41]), "r"(r[42]), "r"(r[43]), "r"(r[44]), "r"(r[45]), "r"(r[46]), "r"(r[47]), "r"(r[48]), "r"(r[49]), "r"(r[50]), "r"(r[51]), "r"(r[52]), "r"(r[53]), "r"(r[54]), "r"(r[55]), "r"(r[56]), "r"(r[57]), "r"(r[58]), "r"(r[59]), "r"(r[60]), "r"(r[61]), "r"(r[62]), "r"(r[63]) : "memory");
}
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t (&r)[64])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63]) : "r"(taddr) : "memory");
}

__global__ void __launch_bounds__(512, 1) k_probe(int rounds, long long *clocks, int *errors)
{
    __shared__ uint32_t tbase_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(&tbase_s)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tbase_s + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * 64);
    uint32_t r[64];
#pragma unroll
    for (int i = 0; i < 64; i++) r[i] = threadIdx.x * 64 + i + blockIdx.x * 7;
    __syncthreads();
    const long long t0 = clock64();
    for (int k = 0; k < rounds; k++) {
        tmem_st64(taddr, r);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 64; i++) r[i] = 0;           // the registers are free here
        __syncthreads();
        tmem_ld64(taddr, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 64; i++) r[i] += 1;
        __syncthreads();
    }
    const long long t1 = clock64();
    int bad = 0;
#pragma unroll
    for (int i = 0; i < 64; i++) bad += r[i] != threadIdx.x * 64 + i + blockIdx.x * 7 + rounds;
    if (bad) atomicAdd(errors, bad);
    if (threadIdx.x == 0) clocks[blockIdx.x] = t1 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase_s), "r"(256u) : "memory");
}

int main()
{
    long long *d_clk; int *d_err;
    cudaMalloc(&d_clk, 148 * sizeof(long long)); cudaMalloc(&d_err, sizeof(int)); cudaMemset(d_err, 0, sizeof(int));
    const int rounds = 1000;
    k_probe<<<148, 512>>>(10, d_clk, d_err);
    cudaMemset(d_err, 0, sizeof(int));
    k_probe<<<148, 512>>>(rounds, d_clk, d_err);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; int err = 0;
    cudaMemcpy(h, d_clk, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(&err, d_err, sizeof(int), cudaMemcpyDeviceToHost);
    long long mx = 0, mn = 1LL << 62;
    for (int i = 0; i < 148; i++) { if (h[i] > mx) mx = h[i]; if (h[i] < mn) mn = h[i]; }
    printf("status %s, errors %d; park + restore of 128 KB per CTA (incl. 2 barriers): %.1f .. %.1f clocks per round trip\n",
           cudaGetErrorString(e), err, (double)mn / rounds, (double)mx / rounds);
    return 0;
}
