// ntt_fp64.cu -- negacyclic NTT / INTT on the FP64 pipe for primes below 2^47 (every CKKS/BGV "scale" prime).
//
// Measured on B200 (tools/ubench): DFMA/DMUL/DADD issue at 0.5 warp-instr/clk/SMSP -- the same rate as IMAD -- on a
// pipe of their own (DFMA + IMAD together reach 0.88). A 64-bit integer modular product costs ~11 FMA-pipe + ~10
// ALU instructions; for q < 2^47 the same product is exact in 5 FP64 operations:
//     h = v*w ; l = fma(v, w, -h) ; t = rint(h / q) ; r = fma(-t, q, h) ; v*w mod q == r + l,  |r + l| < q
// (h is an integer-valued double, l the exact rounding error, t is off by at most one because |h/q| * 2^-53 < 0.15
// for |v| < 24q, so h - t*q is a small integer and exact). Residues travel as integer-valued doubles in a signed lazy
// range: a CT stage adds < q to the magnitude, so the forward transform needs NO intermediate correction; the GS
// inverse doubles the sum side each stage and is renormalised once per radix-16 round. Inputs / outputs in HBM stay
// uint64 (magic-number conversions, exact below 2^52); only the intermediate between the two passes is stored as raw
// doubles. Outputs are canonical, hence bit-identical to the reference's NTTStandard / INTTStandard
// (ring/ntt.go:174-206). Eligibility per prime: LimbConst.fp_ok.
#include <cstdlib>
#include "../../include/lattigo_b200.h"
#include "engine.h"
#include "modarith.cuh"
#include "ntt_arith.cuh"

#ifndef NTT_FP_REG_OUT
#define NTT_FP_REG_OUT 0
#endif

namespace lgpu {

struct FpParams {
    const LimbConst* limbs;
    RowMap rm;
    const u64* in;
    u64* out;
    size_t in_rs, in_bs, out_rs, out_bs;
    int logN;
};

// ---- strided pass ------------------------------------------------------------------------------------------
template <int RL, bool INVERSE>
__global__ void __launch_bounds__(256) ntt_fp_strided_kernel(FpParams p) {
    constexpr int R = 1 << RL;
    const int b = blockIdx.z;
    const LimbConst L = p.limbs[p.rm.limb[blockIdx.y]];
    const int row = p.rm.drow[blockIdx.y];
    const int N = 1 << p.logN;
    const int stride = N >> RL;
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= stride) return;
    const u64* in = p.in + (size_t)b * p.in_bs + (size_t)row * p.in_rs;
    u64* out = p.out + (size_t)b * p.out_bs + (size_t)row * p.out_rs;
    const double q = L.fq, qinv = L.fqinv;
    double x[R];
    if constexpr (!INVERSE) {
#pragma unroll
        for (int k = 0; k < R; k++) x[k] = u2d_any(in[(size_t)k * stride + l], L.q, L.bred_hi);
        const double* tw = L.ftw_fwd;
#pragma unroll
        for (int u = 0; u < RL; u++) {
            const int half = 1 << (RL - 1 - u);
#pragma unroll
            for (int k = 0; k < R; k++) {
                if (k & half) continue;
                fp_fwd_bfly(x[k], x[k + half], __ldg(tw + (1 << u) + (k >> (RL - u))), q, qinv);
            }
        }
#pragma unroll
        for (int k = 0; k < R; k++) out[(size_t)k * stride + l] = (u64)__double_as_longlong(x[k]);   // raw doubles
    } else {
        // input: raw doubles from the chunk pass, renormalised to |x| < 0.66q
#pragma unroll
        for (int k = 0; k < R; k++) x[k] = __longlong_as_double((long long)in[(size_t)k * stride + l]);
        const double* tw = L.ftw_bwd;
#pragma unroll
        for (int u = RL - 1; u >= 1; u--) {
            const int half = 1 << (RL - 1 - u);
#pragma unroll
            for (int k = 0; k < R; k++) {
                if (k & half) continue;
                fp_inv_bfly(x[k], x[k + half], __ldg(tw + (1 << u) + (k >> (RL - u))), q, qinv);
            }
        }
        {   // last stage with N^-1 folded in; |sums| < 2^RL q here
            constexpr int half = 1 << (RL - 1);
#pragma unroll
            for (int k = 0; k < half; k++) {
                const double u = x[k], v = x[k + half];
                const double a = fp_mulmod(__dadd_rn(u, v), L.fninv, q, qinv);
                const double c = fp_mulmod(__dadd_rn(u, -v), L.flast_inv, q, qinv);
                out[(size_t)k * stride + l] = d2u(a < 0.0 ? a + q : a);
                out[(size_t)(k + half) * stride + l] = d2u(c < 0.0 ? c + q : c);
            }
        }
    }
}

// ---- chunk pass ------------------------------------------------------------------------------------------
// Forward chunk pass. One CTA owns a (limb, chunk) pair and walks over `bpc` batch elements: the twiddles of that
// pair stay in L1, and the next element's 16 values per thread are prefetched into registers while the current
// element is in its second and third rounds, so the round-one global-load latency is hidden.
template <int CL>
__global__ void __launch_bounds__((1 << CL) / 16, 2) ntt_fp_chunk_fwd_kernel(FpParams p, int batch, int bpc) {
    constexpr int T = (1 << CL) / 16;
    constexpr int R0 = fp_round_bits(CL, 0), R1 = fp_round_bits(CL, 1), R2 = fp_round_bits(CL, 2);
    static_assert(R0 == 4, "first round is radix-16");
    extern __shared__ double fsm[];
    const int chunk = blockIdx.x, tid = threadIdx.x;
    const LimbConst L = p.limbs[p.rm.limb[blockIdx.y]];
    const int row = p.rm.drow[blockIdx.y];
    const int s1 = p.logN - CL;
    const int b0 = blockIdx.z * bpc;
    const int b1 = min(b0 + bpc, batch);
    const double q = L.fq, qinv = L.fqinv;
    const double* tw = L.ftw_fwd;
    {   // pull the last round's twiddle lines (8 B x 15 per thread) into L1 early
#pragma unroll
        for (int u = 0; u < R2; u++) {
            const int s = s1 + R0 + R1 + u;
            asm volatile("prefetch.global.L1 [%0];" ::"l"(tw + (1 << s) + (chunk << (R0 + R1 + u)) + (tid << u)));
        }
    }
    auto src_of = [&](int b) {
        return (s1 > 0 ? (const u64*)p.out + (size_t)b * p.out_bs + (size_t)row * p.out_rs
                       : p.in + (size_t)b * p.in_bs + (size_t)row * p.in_rs) + ((size_t)chunk << CL);
    };
    u64 raw[16];
    {
        const u64* src = src_of(b0);
#pragma unroll
        for (int k = 0; k < 16; k++) raw[k] = src[k * T + tid];
    }
    for (int b = b0; b < b1; b++) {
        u64* dst = p.out + (size_t)b * p.out_bs + (size_t)row * p.out_rs + ((size_t)chunk << CL);
        {   // round 1 from registers: element k of this thread sits at k*T + tid (hi = 0, lo = tid)
            double x[16];
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = (s1 > 0) ? __longlong_as_double((long long)raw[k]) : u2d_any(raw[k], L.q, L.bred_hi);
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int half = 1 << (3 - u);
                const int twbase = (1 << (s1 + u)) + (chunk << u);
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    if (k & half) continue;
                    fp_fwd_bfly(x[k], x[k + half], __ldg(tw + twbase + (k >> (4 - u))), q, qinv);
                }
            }
#pragma unroll
            for (int k = 0; k < 16; k++) fsm[fpad(k * T + tid)] = x[k];
        }
        double t2[15];
        fp_load_tw<CL, R0, R1>(t2, tw, s1, chunk, tid);
        __syncthreads();
        if (b + 1 < b1) {   // prefetch the next element while rounds 2 and 3 run
            const u64* src = src_of(b + 1);
#pragma unroll
            for (int k = 0; k < 16; k++) raw[k] = src[k * T + tid];
        }
        fp_fwd_round_tw<CL, R0, R1>(fsm, t2, q, qinv, tid);
        double t3[15];
        fp_load_tw<CL, R0 + R1, R2>(t3, tw, s1, chunk, tid);
        __syncthreads();
#if NTT_FP_REG_OUT
        fp_fwd_round_tw_out<CL, R0 + R1, R2>(fsm, t3, q, qinv, tid, dst);   // 16 consecutive coefficients per thread
        __syncthreads();                                                   // tile reads done before the next element's stores
#else
        // copy-out through the tile: every store instruction covers 32 consecutive words (measured 7 % faster than
        // writing the last round's 16 consecutive coefficients per thread straight from registers)
        fp_fwd_round_tw<CL, R0 + R1, R2>(fsm, t3, q, qinv, tid);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const int idx = k * T + tid;
            dst[idx] = fp_canon(fsm[fpad(idx)], q, qinv);
        }
        __syncthreads();
#endif
    }
}

template <int CL>
__global__ void __launch_bounds__((1 << CL) / 16, 2) ntt_fp_chunk_inv_kernel(FpParams p) {
    constexpr int T = (1 << CL) / 16;
    constexpr int R0 = fp_round_bits(CL, 0), R1 = fp_round_bits(CL, 1), R2 = fp_round_bits(CL, 2);
    extern __shared__ double fsm[];
    const int b = blockIdx.z, chunk = blockIdx.x, tid = threadIdx.x;
    const LimbConst L = p.limbs[p.rm.limb[blockIdx.y]];
    const int row = p.rm.drow[blockIdx.y];
    const int s1 = p.logN - CL;
    const u64* src = p.in + (size_t)b * p.in_bs + (size_t)row * p.in_rs + ((size_t)chunk << CL);
    u64* dst = p.out + (size_t)b * p.out_bs + (size_t)row * p.out_rs + ((size_t)chunk << CL);
    const double q = L.fq, qinv = L.fqinv;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const int idx = k * T + tid;
        // inputs are residues < 2q (canonical or lazy); centre them once so that every later bound holds
        fsm[fpad(idx)] = fp_reduce(u2d_any(src[idx], L.q, L.bred_hi), q, qinv);
    }
    __syncthreads();
    fp_inv_round<CL, R0 + R1, R2, 0>(fsm, nullptr, L, s1, chunk, tid);
    __syncthreads();
    fp_inv_round<CL, R0, R1, 0>(fsm, nullptr, L, s1, chunk, tid);
    __syncthreads();
    if (s1 == 0) fp_inv_round<CL, 0, R0, 2>(fsm, dst, L, s1, chunk, tid);
    else         fp_inv_round<CL, 0, R0, 1>(fsm, dst, L, s1, chunk, tid);
}

// ---- launchers ------------------------------------------------------------------------------------------
// batch elements per CTA of the forward chunk pass: as many as possible while keeping >= ~4 waves of CTAs
static int fp_batch_per_cta(int chunks, int rows, int batch) {
    static const int forced = [] { const char* e = getenv("LGPU_NTT_BPC"); return e ? atoi(e) : 0; }();
    if (forced > 0) return forced < batch ? forced : batch;
    int bpc = batch;
    while (bpc > 1 && (long)chunks * rows * ((batch + bpc - 1) / bpc) < 4L * 296) bpc = (bpc + 1) / 2;
    return bpc;
}

template <int CL>
static int fp_launch_chunk(bool inverse, const FpParams& p, dim3 grid, cudaStream_t st) {
    constexpr int C = 1 << CL;
    const size_t smem = (size_t)(C + (C >> 4) + 1) * sizeof(double);
    if (inverse) ntt_fp_chunk_inv_kernel<CL><<<grid, C / 16, smem, st>>>(p);
    else {
        const int batch = grid.z;
        const int bpc = fp_batch_per_cta(grid.x, grid.y, batch);
        dim3 g2(grid.x, grid.y, (batch + bpc - 1) / bpc);
        ntt_fp_chunk_fwd_kernel<CL><<<g2, C / 16, smem, st>>>(p, batch, bpc);
    }
    LGPU_CUDA_OK(cudaGetLastError());
    return 0;
}
template <bool INV>
static int fp_launch_strided(int rl, const FpParams& p, int rows, int batch, cudaStream_t st) {
    const int threads_total = (1 << p.logN) >> rl;
    const int bs = threads_total < 256 ? threads_total : 256;
    dim3 grid((threads_total + bs - 1) / bs, rows, batch);
    switch (rl) {
        case 1: ntt_fp_strided_kernel<1, INV><<<grid, bs, 0, st>>>(p); break;
        case 2: ntt_fp_strided_kernel<2, INV><<<grid, bs, 0, st>>>(p); break;
        case 3: ntt_fp_strided_kernel<3, INV><<<grid, bs, 0, st>>>(p); break;
        case 4: ntt_fp_strided_kernel<4, INV><<<grid, bs, 0, st>>>(p); break;
        case 5: ntt_fp_strided_kernel<5, INV><<<grid, bs, 0, st>>>(p); break;
        default: set_error("unsupported strided radix"); return -1;
    }
    LGPU_CUDA_OK(cudaGetLastError());
    return 0;
}

bool fp64_ntt_supported(const Ctx* c) {
    static const int off = [] { const char* e = getenv("LGPU_NO_FP64_NTT"); return e && atoi(e) ? 1 : 0; }();
    return !off && c->ring_type == 0 && c->logN >= 10 && c->logN <= 17;
}

// rows must all be fp_ok limbs
int launch_ntt_fp64(const Ctx* c, const RowMap& rm, bool inverse, CSpan in, Span out, int batch, cudaStream_t st) {
    FpParams p;
    p.limbs = c->d_limbs; p.rm = rm; p.in = in.p; p.out = out.p;
    p.in_rs = in.row_stride; p.in_bs = in.batch_stride; p.out_rs = out.row_stride; p.out_bs = out.batch_stride;
    p.logN = c->logN;
    const int cl = c->logN > 12 ? 12 : c->logN;
    const int s1 = c->logN - cl;
    dim3 grid(1u << s1, rm.nrows, batch);
    auto chunk = [&](bool inv, const FpParams& q) {
        switch (cl) {
            case 10: return fp_launch_chunk<10>(inv, q, grid, st);
            case 11: return fp_launch_chunk<11>(inv, q, grid, st);
            default: return fp_launch_chunk<12>(inv, q, grid, st);
        }
    };
    if (!inverse) {
        if (s1 > 0 && fp_launch_strided<false>(s1, p, rm.nrows, batch, st)) return -1;
        return chunk(false, p);
    }
    if (chunk(true, p)) return -1;
    if (s1 > 0) {
        FpParams p2 = p;
        p2.in = out.p; p2.in_rs = out.row_stride; p2.in_bs = out.batch_stride;
        return fp_launch_strided<true>(s1, p2, rm.nrows, batch, st);
    }
    return 0;
}

}  // namespace lgpu
