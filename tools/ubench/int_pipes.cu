// Integer pipe throughput micro-benchmark (sm_100a): measured issue rates (warp-instructions / clk / SMSP) of
// IMAD.WIDE.U32 (with and without 64-bit addend), IMAD (lo), LOP3 and mixes, for 1..8 warps per SMSP.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;

#define REP8(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)

template <int MODE>
__global__ void k(u64* out, u32 a, u64 c, int iters, long long* cyc) {
    u64 x[8]; u32 y[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = threadIdx.x + i; y[i] = threadIdx.x * 3 + i; }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#define WIDE(i) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(x[i]) : "r"((u32)x[i]), "r"(a));
#define WIDEC(i) asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(x[i]) : "r"((u32)x[i]), "r"(a), "l"(c));
#define LO(i) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[i]) : "r"(a), "r"((u32)c));
#define LOP(i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[i]) : "r"(a), "r"((u32)c + i));
            if (MODE == 0) { REP8(WIDE) REP8(WIDE) }
            if (MODE == 1) { REP8(WIDEC) REP8(WIDEC) }
            if (MODE == 2) { REP8(LO) REP8(LO) }
            if (MODE == 3) { REP8(LOP) REP8(LOP) }
            if (MODE == 4) { REP8(WIDE) REP8(LOP) }
            if (MODE == 5) { REP8(WIDE) REP8(LO) }
            if (MODE == 6) { REP8(LO) REP8(LOP) }
            if (MODE == 7) { REP8(WIDE) REP8(LOP) REP8(LOP) }
        }
    }
    long long t1 = clock64();
    u64 acc = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) acc ^= x[i] ^ y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, int per_iter, int w) {
    u64* out; long long* cyc; long long h;
    const int threads = 32 * 4 * w;
    cudaMalloc(&out, 148 * threads * 8); cudaMalloc(&cyc, 8);
    const int iters = 4000;
    k<MODE><<<148, threads>>>(out, 3, 5, 10, cyc);
    k<MODE><<<148, threads>>>(out, 3, 5, iters, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-26s warps/SMSP=%d  warp-instr/clk/SMSP=%.3f\n", name, w, (double)iters * per_iter * w / (double)h);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int w : {1, 2, 4, 8}) {
        run<0>("IMAD.WIDE (no addend)", 64, w);
        run<1>("IMAD.WIDE (64b addend)", 64, w);
        run<2>("IMAD lo", 64, w);
        run<3>("LOP3", 64, w);
        run<4>("WIDE + LOP3 1:1", 64, w);
        run<5>("WIDE + IMAD 1:1", 64, w);
        run<6>("IMAD + LOP3 1:1", 64, w);
        run<7>("WIDE + 2 LOP3", 96, w);
    }
    return 0;
}
