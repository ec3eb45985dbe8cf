// Micro-benchmark: sustained FP64 FMA rate of the vector pipe vs DMMA (development aid for the roofline).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma_kernel(double* out, int iters) {
    double a0 = threadIdx.x, a1 = 1.0, a2 = 2.0, a3 = 3.0, a4 = 4.0, a5 = 5.0, a6 = 6.0, a7 = 7.0;
    const double b = 1.0000001, c = 0.9999999;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void dmma_kernel(double* out, int iters) {
    double c0 = 0, c1 = 0, d0 = 0, d1 = 0, e0 = 0, e1 = 0, f0 = 0, f1 = 0;
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
    for (int i = 0; i < iters; ++i) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(e0), "+d"(e1) : "d"(a), "d"(b));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(f0), "+d"(f1) : "d"(a), "d"(b));
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1 + d0 + d1 + e0 + e1 + f0 + f1;
}
// Both kinds interleaved in every warp: if DMMA and DFMA were independent pipes the combined rate would approach the sum.
__global__ void mixed_kernel(double* out, int iters) {
    double c0 = 0, c1 = 0, d0 = 0, d1 = 0;
    double a0 = threadIdx.x, a1 = 1.0, a2 = 2.0, a3 = 3.0, a4 = 4.0, a5 = 5.0, a6 = 6.0, a7 = 7.0;
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
    const double bb = 1.0000001, cc = 0.9999999;
    for (int i = 0; i < iters; ++i) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
        a0 = fma(a0, bb, cc); a1 = fma(a1, bb, cc); a2 = fma(a2, bb, cc); a3 = fma(a3, bb, cc);
        a4 = fma(a4, bb, cc); a5 = fma(a5, bb, cc); a6 = fma(a6, bb, cc); a7 = fma(a7, bb, cc);
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
        a0 = fma(a0, bb, cc); a1 = fma(a1, bb, cc); a2 = fma(a2, bb, cc); a3 = fma(a3, bb, cc);
        a4 = fma(a4, bb, cc); a5 = fma(a5, bb, cc); a6 = fma(a6, bb, cc); a7 = fma(a7, bb, cc);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1 + d0 + d1 + a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
// throughput of 64-bit warp shuffles (two SHFL.BFLY each) next to nothing else
__global__ void shfl_kernel(double* out, int iters) {
    double a0 = threadIdx.x, a1 = 1.0, a2 = 2.0, a3 = 3.0;
    for (int i = 0; i < iters; ++i) {
        a0 = __shfl_xor_sync(0xffffffffu, a0, 1); a1 = __shfl_xor_sync(0xffffffffu, a1, 2);
        a2 = __shfl_xor_sync(0xffffffffu, a2, 4); a3 = __shfl_xor_sync(0xffffffffu, a3, 8);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
}
// latency: one dependent chain per thread, one warp per SM sub-partition
__global__ void dfma_latency(double* out, int iters, long long* clk) {
    double a = threadIdx.x;
    const double b = 1.0000001, c = 0.9999999;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) { a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); }
    long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) *clk = t1 - t0;
}
int main() {
    double* out; long long* clk;
    cudaMalloc(&out, 148 * 64 * 1024 * sizeof(double)); cudaMalloc(&clk, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int threads : {256, 512, 1024}) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0); dfma_kernel<<<148 * 2, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep) printf("DFMA  %4d thr x 296 CTAs: %.2f TFLOP/s\n", threads, 2.0 * 8 * iters * 296.0 * threads / (ms * 1e-3) / 1e12);
            cudaEventRecord(e0); dmma_kernel<<<148 * 2, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep) printf("DMMA  %4d thr x 296 CTAs: %.2f TFLOP/s\n", threads, 2.0 * 256 * 4 * iters * 296.0 * (threads / 32) / (ms * 1e-3) / 1e12);
        }
    }
    for (int threads : {256, 512, 1024}) {
        float ms;
        mixed_kernel<<<148 * 2, threads>>>(out, iters);
        cudaEventRecord(e0); mixed_kernel<<<148 * 2, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        const double dfma = 2.0 * 16 * iters * 296.0 * threads, dmma = 2.0 * 256 * 2 * iters * 296.0 * (threads / 32);
        printf("MIXED %4d thr x 296 CTAs: DFMA %.2f + DMMA %.2f = %.2f TFLOP/s\n", threads, dfma / (ms * 1e-3) / 1e12,
               dmma / (ms * 1e-3) / 1e12, (dfma + dmma) / (ms * 1e-3) / 1e12);
        shfl_kernel<<<148 * 2, threads>>>(out, iters);
        cudaEventRecord(e0); shfl_kernel<<<148 * 2, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("SHFL64 %4d thr x 296 CTAs: %.1f 32-bit warp shuffles / clock / SM (at 1.965 GHz)\n", threads,
               8.0 * iters * 296.0 * (threads / 32) / (ms * 1e-3) / 148.0 / 1.965e9);
    }
    dfma_latency<<<1, 32>>>(out, 10000, clk); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    printf("dependent DFMA latency: %.1f clocks\n", (double)h / 40000.0);
    return 0;
}
