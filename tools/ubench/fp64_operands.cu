// FP64 pipe with realistic operand patterns: are 3 distinct 64-bit source operands per DFMA still 0.5/clk/SMSP?
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
template <int MODE>
__global__ void k(double* out, double a, double c, int iters, long long* cyc) {
    double x[8], y[8], z[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = threadIdx.x + i; y[i] = 1.0 + 1e-9 * (threadIdx.x + i); z[i] = 0.5 * i + threadIdx.x; }
    const double q = 35184372088833.0, qinv = 1.0 / q, M = 6755399441055744.0;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 2; r++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (MODE == 0) { x[i] = __fma_rn(x[i], a, c); x[i] = __fma_rn(x[i], a, c); x[i] = __fma_rn(x[i], a, c); x[i] = __fma_rn(x[i], a, c); }
                if (MODE == 1) { x[i] = __fma_rn(x[i], y[i], z[i]); y[i] = __fma_rn(y[i], z[i], x[i]); z[i] = __fma_rn(z[i], x[i], y[i]); x[i] = __fma_rn(x[i], y[i], z[i]); }
                if (MODE == 2) {  // the 8-op butterfly on (x[i], y[i]) with twiddle z[i]
                    const double h = __dmul_rn(y[i], z[i]);
                    const double l = __fma_rn(y[i], z[i], -h);
                    const double t = __dadd_rn(__fma_rn(h, qinv, M), -M);
                    const double rr = __fma_rn(-t, q, h);
                    const double v = __dadd_rn(rr, l);
                    const double u = x[i];
                    x[i] = __dadd_rn(u, v); y[i] = __dadd_rn(u, -v);
                }
            }
        }
    }
    long long t1 = clock64();
    double acc = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) acc += x[i] + y[i] + z[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE>
void run(const char* name, int per_iter, int w) {
    double* out; long long* cyc; long long h;
    const int threads = 32 * 4 * w;
    cudaMalloc(&out, 148 * threads * 8); cudaMalloc(&cyc, 8);
    const int iters = 4000;
    k<MODE><<<148, threads>>>(out, 1.0000001, 0.5, 10, cyc);
    k<MODE><<<148, threads>>>(out, 1.0000001, 0.5, iters, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-34s warps/SMSP=%d  FP64 warp-instr/clk/SMSP=%.3f\n", name, w, (double)iters * per_iter * w / (double)h);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int w : {1, 2, 4, 8}) {
        run<0>("DFMA x = x*a + c (reused operands)", 64, w);
        run<1>("DFMA 3 distinct register operands", 64, w);
        run<2>("FP64 butterfly (8 ops)", 128, w);
    }
    return 0;
}
