// ntt.cu -- forward / inverse negacyclic NTT over (limb, coeff) row-major uint64 polynomials.
//
// Replaces ring.Ring.NTT / NTTLazy / INTT / INTTLazy and the per-row SubRing forms
// (reference: ring/ntt.go:127-152,174-206, hot loops nttUnrolled16Lazy :258-552 and
// inttLazyUnrolled16 :608-714).
//
// Decomposition (N = 2^logN, one launch covers rows x batch):
//   forward : [strided pass: first s1 = max(0, logN-12) stages, radix-2^s1 in registers, no smem]
//             -> [chunk pass: remaining c = logN - s1 <= 12 stages on contiguous 2^c-element chunks,
//                 radix-16 register rounds exchanged through padded shared memory]
//   inverse : chunk pass (stages in reverse order, Gentleman-Sande) -> strided pass (+ N^-1 scaling).
// Every element crosses HBM twice per transform when logN > 12 (the intermediate stays L2-resident for
// working sets below ~100 MB), once when logN <= 12.
//
// Arithmetic, two variants selected at compile time (template parameter FAST):
//   FAST = false : the reference's lazy Montgomery butterflies (ring/ntt.go:155-171) with the reference's own
//                  4q-correction schedule, so NTT_EXACT_LAZY reproduces NTTLazy's representative in [0, 6q)
//                  bit for bit (used only when a caller asks for the Lazy API).
//   FAST = true  : free-form path for canonical outputs (NTT / INTT / INTTLazy for N >= 16): Shoup products with
//                  precomputed quotients {w, floor(w 2^64 / q)} -- 6 IMAD.WIDE + 4 IMAD per modular product instead
//                  of ~17 IMAD-class instructions for MRedLazy -- Harvey-style lazy ranges whose correction schedule
//                  is computed per prime (none at all for primes below 2^58), the forward output canonicalised by one
//                  Barrett pass in the store epilogue, the inverse's N^-1 folded into its last stage. The canonical
//                  residue is unique, so the result is bit-identical to the reference's NTTStandard / INTTStandard.
#include <cstdlib>
#include "../../include/lattigo_b200.h"
#include "engine.h"
#include "modarith.cuh"
#include "ntt_arith.cuh"

namespace lgpu {

struct NttParams {
    const LimbConst* limbs;
    RowMap rm;
    const u64* in;
    u64* out;
    size_t in_rs, in_bs, out_rs, out_bs;
    int logN;
    int mode;     // NttMode
    int ci;       // conjugate-invariant ring (stage numbering differs) -- not handled by these kernels
};

// ---------------------------------------------------------------------------------------------------
// strided pass: global stages [0, RL). Thread l handles elements {k * (N >> RL) + l}.
// ---------------------------------------------------------------------------------------------------
template <int RL, bool INVERSE, int FAST>
__global__ void __launch_bounds__(256) ntt_strided_kernel(NttParams p) {
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
    const u64 q = L.q, qinv = L.qinv;
    u64 x[R];
#pragma unroll
    for (int k = 0; k < R; k++) x[k] = in[(size_t)k * stride + l];
    if constexpr (!INVERSE && !FAST) {
        const u64* roots = L.roots_fwd;
#pragma unroll
        for (int u = 0; u < RL; u++) {
            const int half = 1 << (RL - 1 - u);
            const bool red = fwd_reduce_flag(u, p.logN);
#pragma unroll
            for (int k = 0; k < R; k++) {
                if (k & half) continue;
                u64 tw = __ldg(roots + (1 << u) + (k >> (RL - u)));
                fwd_bfly(x[k], x[k + half], tw, q, qinv, red);
            }
        }
    } else if constexpr (!INVERSE && FAST) {
        const ulonglong2* tw = L.tw_fwd;
        const u64 nq = 0ull - q, twoq = q << 1;
#pragma unroll
        for (int u = 0; u < RL; u++) {
            const int half = 1 << (RL - 1 - u);
            const bool corr = (FAST == 1) && ((L.fwd_mask >> u) & 1u);
#pragma unroll
            for (int k = 0; k < R; k++) {
                if (k & half) continue;
                const ulonglong2 w = __ldg(tw + (1 << u) + (k >> (RL - u)));
                fast_fwd_bfly(x[k], x[k + half], w, nq, twoq, L.kq, corr);
            }
        }
    } else if constexpr (INVERSE && !FAST) {
        const u64* roots = L.roots_bwd;
#pragma unroll
        for (int u = RL - 1; u >= 0; u--) {
            const int half = 1 << (RL - 1 - u);
#pragma unroll
            for (int k = 0; k < R; k++) {
                if (k & half) continue;
                u64 tw = __ldg(roots + (1 << u) + (k >> (RL - u)));
                inv_bfly(x[k], x[k + half], tw, q, qinv);
            }
        }
        // x N^-1 (Montgomery), canonical: ring/ntt.go:185-206 (mulscalarmontgomeryvec, also for the Lazy API when N >= 16)
#pragma unroll
        for (int k = 0; k < R; k++) x[k] = mred(x[k], L.ninv, q, qinv);
    } else {
        const ulonglong2* tw = L.tw_bwd;
        const bool lazy = (FAST == 2) || (FAST == 1 && L.inv_lazy != 0);
        const int cl = p.logN - RL;   // stages already done by the chunk pass
#pragma unroll
        for (int u = RL - 1; u >= 1; u--) {
            const int half = 1 << (RL - 1 - u);
            // stage index counted from the first inverse stage: cl + (RL-1-u); inputs are < 2^(idx+1) q when lazy
            const u64 addq = lazy ? (q << (cl + (RL - 1 - u) + 1)) : (q << 1);
#pragma unroll
            for (int k = 0; k < R; k++) {
                if (k & half) continue;
                const ulonglong2 w = __ldg(tw + (1 << u) + (k >> (RL - u)));
                fast_inv_bfly(x[k], x[k + half], w, q, addq, !lazy);
            }
        }
        {   // last stage (u = 0) with N^-1 folded in: X = (U + V) * N^-1, Y = (U - V) * (w * N^-1); canonical outputs
            constexpr int half = 1 << (RL - 1);
            const u64 addq = lazy ? (q << p.logN) : (q << 1);
#pragma unroll
            for (int k = 0; k < half; k++) {
                const u64 U = x[k], V = x[k + half];
                const u64 a = shoup_mul(U + V, L.ninv_s, q);
                const u64 c = shoup_mul(U - V + addq, L.last_inv_s, q);
                x[k] = a >= q ? a - q : a;
                x[k + half] = c >= q ? c - q : c;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < R; k++) out[(size_t)k * stride + l] = x[k];
}

// ---------------------------------------------------------------------------------------------------
// chunk pass: global stages [s1, logN) on a contiguous chunk of C = 2^CL elements held in shared memory.
// ---------------------------------------------------------------------------------------------------
template <int CL, int FAST, int MINB = 1>
__global__ void __launch_bounds__((1 << CL) >= 16 ? ((1 << CL) / 16) : 1, MINB)
ntt_chunk_fwd_kernel(NttParams p) {
    constexpr int C = 1 << CL;
    constexpr int T = C / 16;
    constexpr int R0 = round_bits(CL, 0), R1 = round_bits(CL, 1), R2 = round_bits(CL, 2);
    extern __shared__ u64 sm[];
    const int b = blockIdx.z, chunk = blockIdx.x, tid = threadIdx.x;
    const LimbConst L = p.limbs[p.rm.limb[blockIdx.y]];
    const int row = p.rm.drow[blockIdx.y];
    const int s1 = p.logN - CL;
    // for logN > 12 the strided pass already moved the data to `out`
    const u64* src = (s1 > 0 ? (const u64*)p.out + (size_t)b * p.out_bs + (size_t)row * p.out_rs
                             : p.in + (size_t)b * p.in_bs + (size_t)row * p.in_rs) + ((size_t)chunk << CL);
    u64* dst = p.out + (size_t)b * p.out_bs + (size_t)row * p.out_rs + ((size_t)chunk << CL);
    const u64 q = L.q;
    if constexpr (FAST != 0 && CL == 12) {
        // the last round reads 15 twiddle pairs per thread (61 KB per CTA): pull their lines into L1 now, while
        // the first two rounds run, so the round does not start on an L2 round trip
        const ulonglong2* tw = L.tw_fwd;
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int s = s1 + 8 + u;
            const ulonglong2* a = tw + (1 << s) + (chunk << (8 + u)) + (tid << u);
            asm volatile("prefetch.global.L1 [%0];" ::"l"(a));
        }
    }
    fwd_round<CL, 0, R0, true, FAST>(sm, src, L, s1, p.logN, chunk, tid);
    __syncthreads();
    if constexpr (R1 > 0) {
        fwd_round<CL, R0, R1, false, FAST>(sm, nullptr, L, s1, p.logN, chunk, tid);
        __syncthreads();
    }
    if constexpr (R2 > 0) {
        fwd_round<CL, R0 + R1, R2, false, FAST>(sm, nullptr, L, s1, p.logN, chunk, tid);
        __syncthreads();
    }
    const bool canon = (FAST != 0) || (p.mode == NTT_CANONICAL);
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const int idx = k * T + tid;
        u64 v = sm[pad_idx(idx)];
        if (canon) v = bred_add(v, q, L.bred_hi);   // reducevec, ring/ntt.go:176
        dst[idx] = v;
    }
}

template <int CL, int FAST>
__global__ void __launch_bounds__((1 << CL) >= 16 ? ((1 << CL) / 16) : 1)
ntt_chunk_inv_kernel(NttParams p) {
    constexpr int C = 1 << CL;
    constexpr int T = C / 16;
    constexpr int NR = num_rounds(CL);
    constexpr int R0 = round_bits(CL, 0), R1 = round_bits(CL, 1), R2 = round_bits(CL, 2);
    extern __shared__ u64 sm[];
    const int b = blockIdx.z, chunk = blockIdx.x, tid = threadIdx.x;
    const LimbConst L = p.limbs[p.rm.limb[blockIdx.y]];
    const int row = p.rm.drow[blockIdx.y];
    const int s1 = p.logN - CL;
    const u64* src = p.in + (size_t)b * p.in_bs + (size_t)row * p.in_rs + ((size_t)chunk << CL);
    u64* dst = p.out + (size_t)b * p.out_bs + (size_t)row * p.out_rs + ((size_t)chunk << CL);
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const int idx = k * T + tid;
        sm[pad_idx(idx)] = src[idx];
    }
    __syncthreads();
    // deepest round first
    if constexpr (NR == 3) {
        inv_round<CL, R0 + R1, R2, false, false, FAST>(sm, nullptr, L, s1, chunk, tid);
        __syncthreads();
    }
    if constexpr (NR >= 2) {
        inv_round<CL, R0, R1, false, false, FAST>(sm, nullptr, L, s1, chunk, tid);
        __syncthreads();
    }
    if (s1 == 0) inv_round<CL, 0, R0, true, true, FAST>(sm, dst, L, s1, chunk, tid);
    else         inv_round<CL, 0, R0, true, false, FAST>(sm, dst, L, s1, chunk, tid);
}

// ---------------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------------
template <int CL>
static int launch_chunk(bool inverse, int fast, const NttParams& p, dim3 grid, cudaStream_t st) {
    constexpr int C = 1 << CL;
    constexpr int T = C >= 16 ? C / 16 : 1;
    size_t smem = (size_t)(C + (C >> 4) + 1) * sizeof(u64);
    if (inverse) {
        if (fast == 2) ntt_chunk_inv_kernel<CL, 2><<<grid, T, smem, st>>>(p);
        else if (fast == 1) ntt_chunk_inv_kernel<CL, 1><<<grid, T, smem, st>>>(p);
        else ntt_chunk_inv_kernel<CL, 0><<<grid, T, smem, st>>>(p);
    } else {
        if (fast == 2) {
            static const int minb = [] { const char* e = getenv("LGPU_NTT_MINB"); return e ? atoi(e) : 2; }();
            if (CL == 12 && minb == 3) ntt_chunk_fwd_kernel<CL, 2, (CL == 12 ? 3 : 1)><<<grid, T, smem, st>>>(p);
            else if (CL == 12 && minb == 4) ntt_chunk_fwd_kernel<CL, 2, (CL == 12 ? 4 : 1)><<<grid, T, smem, st>>>(p);
            else ntt_chunk_fwd_kernel<CL, 2, (CL == 12 ? 2 : 1)><<<grid, T, smem, st>>>(p);
        }
        else if (fast == 1) ntt_chunk_fwd_kernel<CL, 1><<<grid, T, smem, st>>>(p);
        else ntt_chunk_fwd_kernel<CL, 0><<<grid, T, smem, st>>>(p);
    }
    LGPU_CUDA_OK(cudaGetLastError());
    return 0;
}

static int launch_chunk_dyn(int cl, bool inverse, int fast, const NttParams& p, dim3 grid, cudaStream_t st) {
    switch (cl) {
        case 4: return launch_chunk<4>(inverse, fast, p, grid, st);
        case 5: return launch_chunk<5>(inverse, fast, p, grid, st);
        case 6: return launch_chunk<6>(inverse, fast, p, grid, st);
        case 7: return launch_chunk<7>(inverse, fast, p, grid, st);
        case 8: return launch_chunk<8>(inverse, fast, p, grid, st);
        case 9: return launch_chunk<9>(inverse, fast, p, grid, st);
        case 10: return launch_chunk<10>(inverse, fast, p, grid, st);
        case 11: return launch_chunk<11>(inverse, fast, p, grid, st);
        case 12: return launch_chunk<12>(inverse, fast, p, grid, st);
    }
    set_error("unsupported chunk size");
    return -1;
}

template <bool INV, int FAST>
static int launch_strided(int rl, const NttParams& p, int rows, int batch, cudaStream_t st) {
    const int N = 1 << p.logN;
    const int threads_total = N >> rl;
    const int bs = threads_total < 256 ? threads_total : 256;
    dim3 grid((threads_total + bs - 1) / bs, rows, batch);
    switch (rl) {
        case 1: ntt_strided_kernel<1, INV, FAST><<<grid, bs, 0, st>>>(p); break;
        case 2: ntt_strided_kernel<2, INV, FAST><<<grid, bs, 0, st>>>(p); break;
        case 3: ntt_strided_kernel<3, INV, FAST><<<grid, bs, 0, st>>>(p); break;
        case 4: ntt_strided_kernel<4, INV, FAST><<<grid, bs, 0, st>>>(p); break;
        case 5: ntt_strided_kernel<5, INV, FAST><<<grid, bs, 0, st>>>(p); break;
        default: set_error("unsupported strided radix"); return -1;
    }
    LGPU_CUDA_OK(cudaGetLastError());
    return 0;
}

static bool persist_worthwhile(int limb_transforms) {
    static const int forced = [] { const char* e = getenv("LGPU_NTT_PERSIST"); return e ? atoi(e) : -1; }();
    return forced == 1 || limb_transforms >= 64;
}

// 2 = every limb of the launch needs no lazy correction at all (forward: fwd_mask == 0; inverse: inv_lazy), so the
// correction code is compiled out; 1 = per-limb schedules evaluated at run time.
static int fast_variant(const Ctx* c, const RowMap& rm, bool inverse) {
    for (int r = 0; r < rm.nrows; r++) {
        const LimbConst& L = c->h_limbs[rm.limb[r]];
        if (inverse ? (L.inv_lazy == 0) : (L.fwd_mask != 0)) return 1;
    }
    return 2;
}

static int check_common(const Ctx* c, const RowMap& rm, int batch) {
    if (c->logN < 4 || c->logN > 17) { set_error("NTT requires 16 <= N <= 2^17"); return -1; }
    if (rm.nrows <= 0 || rm.nrows > kMaxRows || batch <= 0 || batch > 65535) { set_error("bad rows/batch"); return -1; }
    return 0;
}

// canonical transforms: rows whose prime qualifies go to the FP64-pipe kernels, the rest to the integer kernels
static bool split_rows_fp64(const Ctx* c, const RowMap& rm, RowMap& fp, RowMap& rest) {
    fp.nrows = rest.nrows = 0;
    if (!fp64_ntt_supported(c)) return false;
    for (int r = 0; r < rm.nrows; r++) {
        RowMap& d = c->h_limbs[rm.limb[r]].fp_ok ? fp : rest;
        d.limb[d.nrows] = rm.limb[r]; d.drow[d.nrows] = rm.drow[r]; d.nrows++;
    }
    return fp.nrows > 0;
}

static int launch_ntt_int(const Ctx* c, const RowMap& rm, CSpan in, Span out, int batch, int mode, cudaStream_t st);
static int launch_intt_int(const Ctx* c, const RowMap& rm, CSpan in, Span out, int batch, int mode, cudaStream_t st);

int launch_ntt(const Ctx* c, const RowMap& rm, CSpan in, Span out, int batch, int mode, cudaStream_t st) {
    if (check_common(c, rm, batch)) return -1;
    if (c->ring_type != 0) return launch_ntt_ci(c, rm, false, in, out, batch, mode == NTT_EXACT_LAZY, st);
    RowMap fp, rest;
    // the tile queue pays a fill / drain of ~D limb-transforms: below ~2 waves of work the two-pass kernels are faster (one polynomial of 44 limbs:
    // 12.4 % vs 10.6 % of the HBM roofline, profiles/r02_ntt_ab.json)
    const bool persist = mode == NTT_CANONICAL && ntt_persist_supported(c, false) && persist_worthwhile(rm.nrows * batch);
    // the two-pass FP64 forward chunk pass stores 128 bits at a time; the persistent kernels use 64-bit accesses only
    const bool vec_ok = persist || (aligned16(out.p) && even_words(out.row_stride, out.batch_stride));
    if (mode == NTT_CANONICAL && vec_ok && split_rows_fp64(c, rm, fp, rest)) {
        {
            ProfScope ps(LGPU_KCLASS_NTT_FWD, st, 16.0 * c->N * fp.nrows * batch, persist ? 1 : (c->logN > 12 ? 2 : 1));
            if (persist ? launch_ntt_persist(c, fp, false, 0, in, out, batch, st) : launch_ntt_fp64(c, fp, false, in, out, batch, st)) return -1;
        }
        if (rest.nrows == 0) return 0;
        return launch_ntt_int(c, rest, in, out, batch, mode, st);
    }
    return launch_ntt_int(c, rm, in, out, batch, mode, st);
}

// Ring.NTT + Ring.MulCoeffsMontgomery in one pass: algorithmic traffic 3 rows (in, other, out) instead of the
// 5 of the two-call sequence.
int launch_ntt_mul_montgomery(const Ctx* c, const RowMap& rm, CSpan in, CSpan other, Span out, int batch, cudaStream_t st) {
    if (check_common(c, rm, batch)) return -1;
    if (!other.p) { set_error("null operand"); return -1; }
    if (other.p == out.p) { set_error("NTT+MulCoeffsMontgomery: the multiplicand cannot be the output"); return -1; }
    if (c->ring_type == 0 && ntt_persist_supported(c, false) && persist_worthwhile(rm.nrows * batch)) {
        RowMap fp, rest;
        if (!split_rows_fp64(c, rm, fp, rest)) { rest = rm; fp.nrows = 0; }
        if (fp.nrows > 0) {
            ProfScope ps(LGPU_KCLASS_NTT_FWD, st, 24.0 * c->N * fp.nrows * batch, 1);
            if (launch_ntt_persist(c, fp, false, 0, in, out, batch, st, other)) return -1;
        }
        // integer rows: their transform is bound by the integer pipes, which the product would share -- measured 0.9x against two launches
        // (profiles/r02_configs.json, C2 q61), so they keep the transform followed by the coefficient-wise kernel
        if (rest.nrows > 0) {
            if (launch_ntt_int(c, rest, in, out, batch, NTT_CANONICAL, st)) return -1;
            return launch_vecop(c, rest, LGPU_OP_MULCOEFFSMONTGOMERY, CSpan{out.p, out.row_stride, out.batch_stride}, other, out, batch, nullptr, nullptr, 0, 0,
                                c->N, st);
        }
        return 0;
    }
    if (launch_ntt(c, rm, in, out, batch, NTT_CANONICAL, st)) return -1;
    return launch_vecop(c, rm, LGPU_OP_MULCOEFFSMONTGOMERY, CSpan{out.p, out.row_stride, out.batch_stride}, other, out, batch, nullptr, nullptr, 0, 0, c->N, st);
}

int launch_intt(const Ctx* c, const RowMap& rm, CSpan in, Span out, int batch, int mode, cudaStream_t st) {
    if (check_common(c, rm, batch)) return -1;
    if (c->ring_type != 0) return launch_ntt_ci(c, rm, true, in, out, batch, mode == NTT_EXACT_LAZY, st);
    RowMap fp, rest;
    if (mode != NTT_REFERENCE_ARITH && split_rows_fp64(c, rm, fp, rest)) {
        {
            const bool persist = ntt_persist_supported(c, true);
            ProfScope ps(LGPU_KCLASS_NTT_INV, st, 16.0 * c->N * fp.nrows * batch, persist ? 1 : (c->logN > 12 ? 2 : 1));
            if (persist ? launch_ntt_persist(c, fp, true, 0, in, out, batch, st) : launch_ntt_fp64(c, fp, true, in, out, batch, st)) return -1;
        }
        if (rest.nrows == 0) return 0;
        return launch_intt_int(c, rest, in, out, batch, mode, st);
    }
    return launch_intt_int(c, rm, in, out, batch, mode, st);
}

static int launch_ntt_int(const Ctx* c, const RowMap& rm, CSpan in, Span out, int batch, int mode, cudaStream_t st) {
    if (check_common(c, rm, batch)) return -1;
    NttParams p;
    p.limbs = c->d_limbs; p.rm = rm; p.in = in.p; p.out = out.p;
    p.in_rs = in.row_stride; p.in_bs = in.batch_stride; p.out_rs = out.row_stride; p.out_bs = out.batch_stride;
    p.logN = c->logN; p.mode = mode; p.ci = 0;
    const int cl = c->logN > 12 ? 12 : c->logN;
    const int s1 = c->logN - cl;
    const int fast = (mode == NTT_CANONICAL) ? fast_variant(c, rm, false) : 0;
    const bool persist = fast != 0 && ntt_persist_supported(c, false) && persist_worthwhile(rm.nrows * batch);
    ProfScope ps(LGPU_KCLASS_NTT_FWD, st, 16.0 * c->N * rm.nrows * batch, (s1 > 0 && !persist) ? 2 : 1);
    if (persist) return launch_ntt_persist(c, rm, false, fast, in, out, batch, st);
    if (s1 > 0) {
        int rc = fast == 2 ? launch_strided<false, 2>(s1, p, rm.nrows, batch, st)
               : fast == 1 ? launch_strided<false, 1>(s1, p, rm.nrows, batch, st)
                           : launch_strided<false, 0>(s1, p, rm.nrows, batch, st);
        if (rc) return -1;
    }
    dim3 grid(1u << s1, rm.nrows, batch);
    return launch_chunk_dyn(cl, false, fast, p, grid, st);
}

static int launch_intt_int(const Ctx* c, const RowMap& rm, CSpan in, Span out, int batch, int mode, cudaStream_t st) {
    // INTTLazy == INTT for N >= 16 (ring/ntt.go:197-206): both are canonical, so both take the fast path;
    // NTT_REFERENCE_ARITH keeps the Montgomery kernels reachable (cross-check in the tests).
    const int fast = (mode != NTT_REFERENCE_ARITH) ? fast_variant(c, rm, true) : 0;
    if (check_common(c, rm, batch)) return -1;
    NttParams p;
    p.limbs = c->d_limbs; p.rm = rm; p.in = in.p; p.out = out.p;
    p.in_rs = in.row_stride; p.in_bs = in.batch_stride; p.out_rs = out.row_stride; p.out_bs = out.batch_stride;
    p.logN = c->logN; p.mode = mode; p.ci = 0;
    const int cl = c->logN > 12 ? 12 : c->logN;
    const int s1 = c->logN - cl;
    const bool persist = fast != 0 && ntt_persist_supported(c, true);
    ProfScope ps(LGPU_KCLASS_NTT_INV, st, 16.0 * c->N * rm.nrows * batch, (s1 > 0 && !persist) ? 2 : 1);
    if (persist) return launch_ntt_persist(c, rm, true, fast, in, out, batch, st);
    dim3 grid(1u << s1, rm.nrows, batch);
    if (launch_chunk_dyn(cl, true, fast, p, grid, st)) return -1;
    if (s1 > 0) {
        // second pass works in place on `out`
        NttParams p2 = p;
        p2.in = out.p; p2.in_rs = out.row_stride; p2.in_bs = out.batch_stride;
        return fast == 2 ? launch_strided<true, 2>(s1, p2, rm.nrows, batch, st)
             : fast == 1 ? launch_strided<true, 1>(s1, p2, rm.nrows, batch, st)
                         : launch_strided<true, 0>(s1, p2, rm.nrows, batch, st);
    }
    return 0;
}

}  // namespace lgpu
