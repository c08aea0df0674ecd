// FP64 pipe throughput micro-benchmark (sm_100a): DFMA / DADD / conversions, alone and mixed with integer work.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;
#define REP8(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)

template <int MODE>
__global__ void k(double* out, double a, double c, u32 ia, int iters, long long* cyc) {
    double x[8]; u32 y[8]; u64 z[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = threadIdx.x + i; y[i] = threadIdx.x * 3 + i; z[i] = threadIdx.x + 5 * i; }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#define DFMA(i) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[i]) : "d"(a), "d"(c));
#define DADD(i) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(x[i]) : "d"(c));
#define DMUL(i) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(x[i]) : "d"(a));
#define LOP(i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[i]) : "r"(ia), "r"(ia + i));
#define LO(i) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[i]) : "r"(ia), "r"(ia + 1));
#define CVTUD(i) asm volatile("cvt.rn.f64.u64 %0, %1;" : "=d"(x[i]) : "l"(z[i])); z[i] += (u64)x[i] == 0;
#define CVTDU(i) asm volatile("cvt.rzi.u64.f64 %0, %1;" : "=l"(z[i]) : "d"(x[i])); 
            if (MODE == 0) { REP8(DFMA) REP8(DFMA) }
            if (MODE == 1) { REP8(DADD) REP8(DADD) }
            if (MODE == 2) { REP8(DMUL) REP8(DMUL) }
            if (MODE == 3) { REP8(DFMA) REP8(LOP) }
            if (MODE == 4) { REP8(DFMA) REP8(LO) }
            if (MODE == 5) { REP8(DFMA) REP8(LO) REP8(LOP) }
            if (MODE == 6) { REP8(CVTDU) }
        }
    }
    long long t1 = clock64();
    double acc = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) acc += x[i] + y[i] + (double)z[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, int per_iter, int w) {
    double* out; long long* cyc; long long h;
    const int threads = 32 * 4 * w;
    cudaMalloc(&out, 148 * threads * 8); cudaMalloc(&cyc, 8);
    const int iters = 4000;
    k<MODE><<<148, threads>>>(out, 1.0000001, 0.5, 3, 10, cyc);
    k<MODE><<<148, threads>>>(out, 1.0000001, 0.5, 3, iters, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-26s warps/SMSP=%d  warp-instr/clk/SMSP=%.3f\n", name, w, (double)iters * per_iter * w / (double)h);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int w : {1, 4, 8}) {
        run<0>("DFMA", 64, w);
        run<1>("DADD", 64, w);
        run<2>("DMUL", 64, w);
        run<3>("DFMA + LOP3 1:1", 64, w);
        run<4>("DFMA + IMAD 1:1", 64, w);
        run<5>("DFMA + IMAD + LOP3", 96, w);
        run<6>("CVT f64->u64", 32, w);
    }
    return 0;
}
