// ntt_ci.cu -- conjugate-invariant transforms, Z[X + X^-1]/(X^2N + 1) (ring.ConjugateInvariant).
//
// Reference: ring/ntt.go:716-1311 (NTTConjugateInvariant[Lazy] :717-726, INTTConjugateInvariant[Lazy] :728-737,
// nttCoreConjugateInvariantLazy :740-1088, inttCoreConjugateInvariantLazy :1090-1311).
//
// Structure (N coefficients, NthRoot = 4N, root tables of 2N entries):
//   forward : fold   p2[j] = p1[j] + 2q - MRedLazy(p1[N-j], roots[1])   (j = 1..N-1, p2[0] = p1[0])
//             then stages s = 0..logN-1 with span t = N >> (s+1), block i = j / 2t, twiddle roots[2^(s+1) + i]
//             -- the stages 1..logN of a 2N-point negacyclic transform restricted to the first half of the array --
//             with the reference's correction schedule (U >= 4q test on stages where bits.Len64(2^(s+1)) is odd <=> s odd).
//   inverse : stages of span t = 1, 2, .., N/2 with twiddle roots[N/t + i], then the unfold with roots[1] of the
//             backward table, p2[0] = CRed(2 p2[0]), then * NInv (MRed, or MRedLazy for the Lazy variant).
// All butterflies are the reference's Montgomery ones, so the Lazy outputs are the same representatives bit for bit.
//
// This path is rare on the hot path (CKKS-real parameters only), so it is written for exactness and a sane number of
// passes, not for the last percent: spans >= 4096 take one global radix-2 pass each, everything below runs inside one
// 4096-element shared-memory chunk kernel.
#include "../../include/lattigo_b200.h"
#include "engine.h"
#include "modarith.cuh"
#include "ntt_arith.cuh"

namespace lgpu {

struct CiParams {
    const LimbConst* limbs;
    RowMap rm;
    const u64* in;
    u64* out;
    size_t in_rs, in_bs, out_rs, out_bs;
    int logN;
    int lazy;
    int s;   // stage (forward) / log2 span (inverse) for the global radix-2 kernels
};

__device__ __forceinline__ bool ci_reduce_flag(int s) { return (s & 1) != 0; }  // ring/ntt.go:853: bits.Len64(m) odd, m = 2^(s+1)

// fold / unfold: one thread per pair (j, N-j); thread 0 also handles j = 0 and j = N/2.
template <bool INVERSE>
__global__ void __launch_bounds__(256) ci_fold_kernel(CiParams p) {
    const LimbConst L = p.limbs[p.rm.limb[blockIdx.y]];
    const int row = p.rm.drow[blockIdx.y];
    const int N = 1 << p.logN, n2 = N >> 1;
    const u64* in = p.in + (size_t)blockIdx.z * p.in_bs + (size_t)row * p.in_rs;
    u64* out = p.out + (size_t)blockIdx.z * p.out_bs + (size_t)row * p.out_rs;
    const u64 q = L.q, qinv = L.qinv, twoq = q << 1;
    const u64 F = INVERSE ? L.roots_bwd[1] : L.roots_fwd[1];
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n2) return;
    if (j == 0) {
        const u64 mid = in[n2];
        u64 m = mid + twoq - mred_lazy(mid, F, q, qinv);
        u64 z = in[0];
        if (INVERSE) {
            z = cred(z << 1, q);                                        // ring/ntt.go:1156
            z = p.lazy ? mred_lazy(z, L.ninv, q, qinv) : mred(z, L.ninv, q, qinv);
            m = p.lazy ? mred_lazy(m, L.ninv, q, qinv) : mred(m, L.ninv, q, qinv);
        }
        out[n2] = m;
        out[0] = z;
        return;
    }
    const u64 x = in[j], y = in[N - j];
    u64 a = x + twoq - mred_lazy(y, F, q, qinv);
    u64 b = y + twoq - mred_lazy(x, F, q, qinv);
    if (INVERSE) {
        a = p.lazy ? mred_lazy(a, L.ninv, q, qinv) : mred(a, L.ninv, q, qinv);
        b = p.lazy ? mred_lazy(b, L.ninv, q, qinv) : mred(b, L.ninv, q, qinv);
    }
    out[j] = a;
    out[N - j] = b;
}

// one global radix-2 stage, in place on `out`. Forward: stage p.s (span t = N >> (s+1)); inverse: span t = 1 << p.s.
template <bool INVERSE>
__global__ void __launch_bounds__(256) ci_stage_kernel(CiParams p) {
    const LimbConst L = p.limbs[p.rm.limb[blockIdx.y]];
    const int row = p.rm.drow[blockIdx.y];
    const int N = 1 << p.logN;
    u64* a = p.out + (size_t)blockIdx.z * p.out_bs + (size_t)row * p.out_rs;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;   // butterfly index
    if (b >= (N >> 1)) return;
    const int lt = INVERSE ? p.s : p.logN - 1 - p.s;       // log2 span
    const int i = b >> lt;                                  // block
    const int j = ((i << 1) << lt) + (b & ((1 << lt) - 1));
    const int m = N >> lt;                                  // = 2^(s+1) forward, N/t inverse
    u64 X = a[j], Y = a[j + (1 << lt)];
    if (INVERSE) inv_bfly(X, Y, L.roots_bwd[m + i], L.q, L.qinv);
    else fwd_bfly(X, Y, L.roots_fwd[m + i], L.q, L.qinv, ci_reduce_flag(p.s));
    a[j] = X;
    a[j + (1 << lt)] = Y;
}

// all stages of span < 2^CL inside one 2^CL-element chunk (CL = min(logN, 12)), 2^(CL-3) threads x 4 butterflies.
template <bool INVERSE>
__global__ void __launch_bounds__(512) ci_chunk_kernel(CiParams p, int CL) {
    extern __shared__ u64 sm[];
    const LimbConst L = p.limbs[p.rm.limb[blockIdx.y]];
    const int row = p.rm.drow[blockIdx.y];
    const int chunk = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
    const int n = 1 << CL, N = 1 << p.logN;
    // forward reads `out` (the fold and the global stages already ran in place there); inverse reads `in`
    const u64* src = INVERSE ? p.in + (size_t)blockIdx.z * p.in_bs + (size_t)row * p.in_rs + ((size_t)chunk << CL)
                             : p.out + (size_t)blockIdx.z * p.out_bs + (size_t)row * p.out_rs + ((size_t)chunk << CL);
    u64* dst = p.out + (size_t)blockIdx.z * p.out_bs + (size_t)row * p.out_rs + ((size_t)chunk << CL);
    const u64 q = L.q, qinv = L.qinv;
    for (int k = tid; k < n; k += T) sm[pad_idx(k)] = src[k];
    __syncthreads();
    const int s1 = p.logN - CL;
    for (int u = 0; u < CL; u++) {
        const int lt = INVERSE ? u : CL - 1 - u;            // log2 span of this stage
        const int s = p.logN - 1 - lt;                      // forward stage number
        const int m = N >> lt;
        for (int b = tid; b < (n >> 1); b += T) {
            const int il = b >> lt;                         // block within the chunk
            const int j = ((il << 1) << lt) + (b & ((1 << lt) - 1));
            const int i = (chunk << (CL - 1 - lt)) + il;    // global block index
            u64 X = sm[pad_idx(j)], Y = sm[pad_idx(j + (1 << lt))];
            if (INVERSE) inv_bfly(X, Y, L.roots_bwd[m + i], q, qinv);
            else fwd_bfly(X, Y, L.roots_fwd[m + i], q, qinv, ci_reduce_flag(s));
            sm[pad_idx(j)] = X;
            sm[pad_idx(j + (1 << lt))] = Y;
        }
        __syncthreads();
    }
    (void)s1;
    for (int k = tid; k < n; k += T) {
        u64 v = sm[pad_idx(k)];
        if (!INVERSE && !p.lazy) v = bred_add(v, q, L.bred_hi);   // NTTConjugateInvariant: reducevec, ring/ntt.go:719
        dst[k] = v;
    }
}

int launch_ntt_ci(const Ctx* c, const RowMap& rm, bool inverse, CSpan in, Span out, int batch, int lazy, cudaStream_t st) {
    CiParams p;
    p.limbs = c->d_limbs; p.rm = rm; p.in = in.p; p.out = out.p;
    p.in_rs = in.row_stride; p.in_bs = in.batch_stride; p.out_rs = out.row_stride; p.out_bs = out.batch_stride;
    p.logN = c->logN; p.lazy = lazy; p.s = 0;
    const int N = c->N;
    const int CL = c->logN > 12 ? 12 : c->logN;
    const int s1 = c->logN - CL;
    const int ct = (1 << CL) / 8 < 32 ? 32 : (1 << CL) / 8;
    const size_t smem = (size_t)((1 << CL) + ((1 << CL) >> 4) + 1) * sizeof(u64);
    const int half = N >> 1;
    const dim3 gpair((unsigned)((half + 255) / 256), rm.nrows, batch);
    const dim3 gchunk(1u << s1, rm.nrows, batch);
    ProfScope ps(inverse ? LGPU_KCLASS_NTT_INV : LGPU_KCLASS_NTT_FWD, st, 16.0 * N * rm.nrows * batch, 2 + s1);
    if (!inverse) {
        ci_fold_kernel<false><<<gpair, 256, 0, st>>>(p);
        for (int s = 0; s < s1; s++) { p.s = s; ci_stage_kernel<false><<<gpair, 256, 0, st>>>(p); }
        ci_chunk_kernel<false><<<gchunk, ct, smem, st>>>(p, CL);
    } else {
        ci_chunk_kernel<true><<<gchunk, ct, smem, st>>>(p, CL);
        for (int lt = CL; lt < c->logN; lt++) { p.s = lt; ci_stage_kernel<true><<<gpair, 256, 0, st>>>(p); }
        CiParams pf = p;
        pf.in = out.p; pf.in_rs = out.row_stride; pf.in_bs = out.batch_stride;
        ci_fold_kernel<true><<<gpair, 256, 0, st>>>(pf);
    }
    LGPU_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace lgpu
