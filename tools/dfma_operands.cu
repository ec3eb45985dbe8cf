// Micro-benchmark: does the FP64 FMA rate depend on how many DISTINCT register operands an instruction reads?
// (a) acc = fma(acc, b, c)  with b, c shared by all chains (operand-reuse friendly, what tools/fp64_peak.cu measures)
// (b) acc_i = fma(x_i, y_i, acc_i)  with 8 distinct (x_i, y_i) pairs: three distinct 64-bit operands per instruction
// (c) the Jacobi rotation body: 8 FMA on two complex numbers with 4 broadcast parameters
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_a(double* out, const double* in, int iters) {
    double a[8];
    for (int i = 0; i < 8; ++i) a[i] = in[threadIdx.x + 32 * i];
    const double b = in[0] + 1.0000001, c = in[1] + 0.9999999;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b, c);
    double s = 0; for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_b(double* out, const double* in, int iters) {
    double a[8], x[8], y[8];
    for (int i = 0; i < 8; ++i) { a[i] = in[threadIdx.x + 32 * i]; x[i] = in[threadIdx.x + 7 * i + 1] * 1e-3; y[i] = in[threadIdx.x + 5 * i + 2] * 1e-3; }
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma(x[i], y[i], a[i]);
    double s = 0; for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_c(double* out, const double* in, int iters) {
    double xr[4], xi[4], yr[4], yi[4];
    for (int i = 0; i < 4; ++i) { xr[i] = in[threadIdx.x + i]; xi[i] = in[threadIdx.x + 4 + i]; yr[i] = in[threadIdx.x + 8 + i]; yi[i] = in[threadIdx.x + 12 + i]; }
    const double tpr = in[3] * 1e-3, tpi = in[4] * 1e-3, sqr = -tpr, sqi = tpi;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const double xor_ = xr[i], xoi = xi[i], yor = yr[i], yoi = yi[i];
            xr[i] = fma(-tpr, yor, fma(tpi, yoi, xor_));
            xi[i] = fma(-tpr, yoi, fma(-tpi, yor, xoi));
            yr[i] = fma(sqr, xor_, fma(-sqi, xoi, yor));
            yi[i] = fma(sqr, xoi, fma(sqi, xor_, yoi));
        }
    double s = 0; for (int i = 0; i < 4; ++i) s += xr[i] + xi[i] + yr[i] + yi[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// (d) the inner-product body: 4 FMA per complex column into 2 accumulators per pair, 4 independent pairs
__global__ void k_d(double* out, const double* in, int iters) {
    double xr[4], xi[4], yr[4], yi[4], cr[4] = {0, 0, 0, 0}, ci[4] = {0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) { xr[i] = in[threadIdx.x + i]; xi[i] = in[threadIdx.x + 4 + i]; yr[i] = in[threadIdx.x + 8 + i]; yi[i] = in[threadIdx.x + 12 + i]; }
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            cr[i] = fma(xr[i], yr[i], fma(xi[i], yi[i], cr[i]));
            ci[i] = fma(xi[i], yr[i], fma(-xr[i], yi[i], ci[i]));
        }
    double s = 0; for (int i = 0; i < 4; ++i) s += cr[i] + ci[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename K>
void run(const char* name, K k, int ctas_per_sm, int threads, double fma_per_thread_iter, double* out, double* in) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    k<<<148 * ctas_per_sm, threads>>>(out, in, iters);
    cudaEventRecord(e0); k<<<148 * ctas_per_sm, threads>>>(out, in, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-28s %2d warps/SM: %6.2f TFLOP/s\n", name, ctas_per_sm * threads / 32,
           2.0 * fma_per_thread_iter * iters * 148.0 * ctas_per_sm * threads / (ms * 1e-3) / 1e12);
}
int main() {
    double *out, *in;
    cudaMalloc(&out, 148 * 8 * 1024 * sizeof(double)); cudaMalloc(&in, 4096 * sizeof(double));
    double h[4096]; for (int i = 0; i < 4096; ++i) h[i] = 1e-3 * (i % 97) + 0.5;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    for (int w : {4, 8, 12, 16, 32}) {
        const int threads = 128, ctas = w / 4;
        run("(a) fma(acc, b, c) shared b,c", k_a, ctas, threads, 8, out, in);
        run("(b) fma(x_i, y_i, acc_i)", k_b, ctas, threads, 8, out, in);
        run("(c) rotation body", k_c, ctas, threads, 32, out, in);
        run("(d) inner-product body", k_d, ctas, threads, 16, out, in);
    }
    return 0;
}
