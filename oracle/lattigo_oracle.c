/*
 * oracle/lattigo_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, gcc) of the arithmetic on Lattigo's RNS ring hot
 * path. It exists to check the CUDA path in lattigo_b200/ and to serve as the
 * "port" CPU baseline in bench.py. Nothing under lattigo_b200/ may link, import
 * or call this file. Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the upstream tree tuneinsight/lattigo v6.2.0). Parity is pinned by
 * tests/test_oracle_golden.py against the golden NTT vectors of
 * ring/ntt_test.go:10-89 (tests/golden/ntt_vectors.json) and by big-integer
 * property tests mirroring ring/ring_test.go.
 *
 * All arithmetic is uint64 wrapping unless stated; the 128-bit products use
 * unsigned __int128 (bits.Mul64 in the reference).
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>

typedef uint64_t u64;
typedef unsigned __int128 u128;

static inline u64 mulhi(u64 a, u64 b) { return (u64)(((u128)a * b) >> 64); }

/* ---- ring/modular_reduction.go ---------------------------------------- */

/* MRed: ring/modular_reduction.go:78-86 */
static inline u64 mred(u64 x, u64 y, u64 q, u64 qinv) {
    u128 m = (u128)x * y;
    u64 mhi = (u64)(m >> 64), mlo = (u64)m;
    u64 hhi = mulhi(mlo * qinv, q);
    u64 r = mhi - hhi + q;
    if (r >= q) r -= q;
    return r;
}
/* MRedLazy: ring/modular_reduction.go:90-95 */
static inline u64 mred_lazy(u64 x, u64 y, u64 q, u64 qinv) {
    u128 m = (u128)x * y;
    u64 mhi = (u64)(m >> 64), mlo = (u64)m;
    u64 hhi = mulhi(mlo * qinv, q);
    return mhi - hhi + q;
}
/* MForm: ring/modular_reduction.go:11-35 ; bred = {hi, lo} of floor(2^128/q) */
static inline u64 mform(u64 a, u64 q, const u64 *bred) {
    u64 mhi = mulhi(a, bred[1]);
    u64 r = (u64)(-(a * bred[0] + mhi)) * q;
    if (r >= q) r -= q;
    return r;
}
/* MFormLazy: ring/modular_reduction.go:40-45 */
static inline u64 mform_lazy(u64 a, u64 q, const u64 *bred) {
    u64 mhi = mulhi(a, bred[1]);
    return (u64)(-(a * bred[0] + mhi)) * q;
}
/* IMForm: ring/modular_reduction.go:49-56 */
static inline u64 imform(u64 a, u64 q, u64 qinv) {
    u64 r = mulhi(a * qinv, q);
    r = q - r;
    if (r >= q) r -= q;
    return r;
}
/* BRedAdd: ring/modular_reduction.go:110-117 */
static inline u64 bred_add(u64 a, u64 q, const u64 *bred) {
    u64 mhi = mulhi(a, bred[0]);
    u64 r = a - mhi * q;
    if (r >= q) r -= q;
    return r;
}
/* BRedAddLazy: ring/modular_reduction.go:121-124 */
static inline u64 bred_add_lazy(u64 a, u64 q, const u64 *bred) {
    return a - mulhi(a, bred[0]) * q;
}
/* BRedLazy: ring/modular_reduction.go:166-196 */
static inline u64 bred_lazy(u64 x, u64 y, u64 q, const u64 *bred) {
    u128 m = (u128)x * y;
    u64 mhi = (u64)(m >> 64), mlo = (u64)m;
    u64 r = mhi * bred[0];
    u128 h = (u128)mlo * bred[0];
    u64 hhi = (u64)(h >> 64), hlo = (u64)h;
    r += hhi;
    u64 lhi = mulhi(mlo, bred[1]);
    u64 s0 = hlo + lhi;
    u64 carry = s0 < hlo;
    r += carry;
    h = (u128)mhi * bred[1];
    hhi = (u64)(h >> 64); hlo = (u64)h;
    r += hhi;
    u64 s1 = hlo + s0;
    carry = s1 < hlo;
    r += carry;
    return mlo - r * q;
}
/* BRed: ring/modular_reduction.go:127-162 */
static inline u64 bred(u64 x, u64 y, u64 q, const u64 *brc) {
    u64 r = bred_lazy(x, y, q, brc);
    if (r >= q) r -= q;
    return r;
}
/* CRed: ring/modular_reduction.go:200-205 */
static inline u64 cred(u64 a, u64 q) { return a >= q ? a - q : a; }

/* exported scalar wrappers (for the edge-operand tests, ring/ring_test.go:537-673) */
u64 lo_mred(u64 x, u64 y, u64 q, u64 qinv) { return mred(x, y, q, qinv); }
u64 lo_mred_lazy(u64 x, u64 y, u64 q, u64 qinv) { return mred_lazy(x, y, q, qinv); }
u64 lo_mform(u64 a, u64 q, const u64 *b) { return mform(a, q, b); }
u64 lo_mform_lazy(u64 a, u64 q, const u64 *b) { return mform_lazy(a, q, b); }
u64 lo_imform(u64 a, u64 q, u64 qinv) { return imform(a, q, qinv); }
u64 lo_bred(u64 x, u64 y, u64 q, const u64 *b) { return bred(x, y, q, b); }
u64 lo_bred_lazy(u64 x, u64 y, u64 q, const u64 *b) { return bred_lazy(x, y, q, b); }
u64 lo_bred_add(u64 a, u64 q, const u64 *b) { return bred_add(a, q, b); }
u64 lo_bred_add_lazy(u64 a, u64 q, const u64 *b) { return bred_add_lazy(a, q, b); }
u64 lo_cred(u64 a, u64 q) { return cred(a, q); }

/* ---- ring/ntt.go ------------------------------------------------------- */

/* butterfly: ring/ntt.go:155-161 (reduce=1) and the reduce=0 inlined form
 * ring/ntt.go:351-383 */
static inline void bfly(u64 *X, u64 *Y, u64 psi, u64 q, u64 qinv, int reduce) {
    u64 U = *X, V = *Y;
    if (reduce && U >= 4 * q) U -= 4 * q;
    V = mred_lazy(V, psi, q, qinv);
    *X = U + V;
    *Y = U + 2 * q - V;
}
/* invbutterfly: ring/ntt.go:164-171 */
static inline void ibfly(u64 *X, u64 *Y, u64 psi, u64 q, u64 qinv) {
    u64 U = *X, V = *Y;
    u64 x = U + V;
    if (x >= 2 * q) x -= 2 * q;
    *X = x;
    *Y = mred_lazy(U + 4 * q - V, psi, q, qinv);
}

/* nttCoreLazy: ring/ntt.go:209-221. Both code paths restated with their exact
 * lazy-reduction schedule:
 *   N < 16  -> nttLazy (ring/ntt.go:223-257): every stage uses butterfly().
 *   N >= 16 -> nttUnrolled16Lazy (ring/ntt.go:258-552): stage m=1 never
 *              reduces U (:275-310); stages 2 <= m < N/2 reduce iff
 *              bits.Len64(m) is odd (:318); the last stage (t==1) always
 *              reduces (:500-518).
 * Output range [0, 6q). */
void lo_ntt_lazy(const u64 *p1, u64 *p2, int N, u64 q, u64 qinv, const u64 *roots) {
    int t = N >> 1;
    int small = N < 16;
    u64 F = roots[1];
    for (int j = 0; j < t; j++) {
        u64 x = p1[j], y = p1[j + t];
        bfly(&x, &y, F, q, qinv, small);
        p2[j] = x; p2[j + t] = y;
    }
    for (int m = 2; m < N; m <<= 1) {
        t >>= 1;
        int len = 0; for (int mm = m; mm; mm >>= 1) len++;
        int reduce = small || (t == 1) || (len & 1);
        for (int i = 0; i < m; i++) {
            int j1 = (i * t) << 1;
            F = roots[m + i];
            for (int j = j1; j < j1 + t; j++) bfly(&p2[j], &p2[j + t], F, q, qinv, reduce);
        }
    }
}
/* NTTStandard: ring/ntt.go:174-177 (nttCoreLazy then reducevec/BRedAdd) */
void lo_ntt(const u64 *p1, u64 *p2, int N, u64 q, u64 qinv, const u64 *brc, const u64 *roots) {
    lo_ntt_lazy(p1, p2, N, q, qinv, roots);
    for (int i = 0; i < N; i++) p2[i] = bred_add(p2[i], q, brc);
}
/* inttCoreLazy: ring/ntt.go:554-714 (inttLazy and inttLazyUnrolled16 perform
 * the same butterflies in the same order of stages). */
static void intt_core(const u64 *p1, u64 *p2, int N, u64 q, u64 qinv, const u64 *roots) {
    int t = 1, h = N >> 1;
    for (int i = 0, j1 = 0; i < h; i++, j1 += 2 * t) {
        u64 F = roots[h + i];
        for (int j = j1; j < j1 + t; j++) {
            u64 x = p1[j], y = p1[j + t];
            ibfly(&x, &y, F, q, qinv);
            p2[j] = x; p2[j + t] = y;
        }
    }
    t <<= 1;
    for (int m = N >> 1; m > 1; m >>= 1) {
        h = m >> 1;
        for (int i = 0, j1 = 0; i < h; i++, j1 += 2 * t) {
            u64 F = roots[h + i];
            for (int j = j1; j < j1 + t; j++) ibfly(&p2[j], &p2[j + t], F, q, qinv);
        }
        t <<= 1;
    }
}
/* INTTStandard: ring/ntt.go:185-194 */
void lo_intt(const u64 *p1, u64 *p2, int N, u64 ninv, u64 q, u64 qinv, const u64 *roots) {
    intt_core(p1, p2, N, q, qinv, roots);
    for (int i = 0; i < N; i++) p2[i] = mred(p2[i], ninv, q, qinv);
}
/* INTTStandardLazy: ring/ntt.go:197-206 -- NB for N >= 16 it calls the
 * NON-lazy mulscalarmontgomeryvec, so the output is fully reduced. */
void lo_intt_lazy(const u64 *p1, u64 *p2, int N, u64 ninv, u64 q, u64 qinv, const u64 *roots) {
    intt_core(p1, p2, N, q, qinv, roots);
    if (N < 16) for (int i = 0; i < N; i++) p2[i] = mred_lazy(p2[i], ninv, q, qinv);
    else        for (int i = 0; i < N; i++) p2[i] = mred(p2[i], ninv, q, qinv);
}

/* ---- Conjugate-invariant NTT, ring/ntt.go:716-1311 -------------------- */
/* nttCoreConjugateInvariantLazy: generic form nttConjugateInvariantLazy
 * (ring/ntt.go:754-786) and unrolled form (:788-1088). Same butterflies; the
 * unrolled form (N >= 16) skips the U >= 4q correction on stages where
 * bits.Len64(m) is even (:853) -- here also for the t==1 stage (:1031-1060). */
void lo_ntt_ci_lazy(const u64 *p1, u64 *p2, int N, u64 q, u64 qinv, const u64 *roots) {
    u64 twoQ = 2 * q;
    int small = N < 16;
    int t = N, n2 = N >> 1;
    u64 F = roots[1];
    for (int jx = 1, jy = N - 1; jx < n2; jx++, jy--) {
        u64 xin = p1[jx], yin = p1[jy];
        p2[jx] = xin + twoQ - mred_lazy(yin, F, q, qinv);
        p2[jy] = yin + twoQ - mred_lazy(xin, F, q, qinv);
    }
    {
        u64 mid = p1[n2];
        p2[n2] = mid + twoQ - mred_lazy(mid, F, q, qinv);
    }
    p2[0] = p1[0];
    for (int m = 2; m < 2 * N; m <<= 1) {
        t >>= 1;
        int h = m >> 1;
        int len = 0; for (int mm = m; mm; mm >>= 1) len++;
        int reduce = small || (len & 1);
        for (int i = 0, j1 = 0; i < h; i++, j1 += 2 * t) {
            F = roots[m + i];
            for (int j = j1; j < j1 + t; j++) bfly(&p2[j], &p2[j + t], F, q, qinv, reduce);
        }
    }
}
/* NTTConjugateInvariant: ring/ntt.go:717-720 */
void lo_ntt_ci(const u64 *p1, u64 *p2, int N, u64 q, u64 qinv, const u64 *brc, const u64 *roots) {
    lo_ntt_ci_lazy(p1, p2, N, q, qinv, roots);
    for (int i = 0; i < N; i++) p2[i] = bred_add(p2[i], q, brc);
}
/* inttCoreConjugateInvariantLazy: ring/ntt.go:1090-1311 (generic form :1104-1158) */
static void intt_ci_core(const u64 *p1, u64 *p2, int N, u64 q, u64 qinv, const u64 *roots) {
    u64 twoQ = 2 * q;
    int t = 1, h = N >> 1, n2 = N >> 1;
    for (int i = 0, j1 = 0; i < h; i++, j1 += 2 * t) {
        u64 F = roots[N + i];
        for (int j = j1; j < j1 + t; j++) {
            u64 x = p1[j], y = p1[j + t];
            ibfly(&x, &y, F, q, qinv);
            p2[j] = x; p2[j + t] = y;
        }
    }
    t <<= 1;
    for (int m = N >> 1; m > 1; m >>= 1) {
        h = m >> 1;
        for (int i = 0, j1 = 0; i < h; i++, j1 += 2 * t) {
            u64 F = roots[m + i];
            for (int j = j1; j < j1 + t; j++) ibfly(&p2[j], &p2[j + t], F, q, qinv);
        }
        t <<= 1;
    }
    u64 F = roots[1];
    for (int jx = 1, jy = N - 1; jx < n2; jx++, jy--) {
        u64 a = p2[jx], b = p2[jy];
        p2[jx] = a + twoQ - mred_lazy(b, F, q, qinv);
        p2[jy] = b + twoQ - mred_lazy(a, F, q, qinv);
    }
    p2[n2] = p2[n2] + twoQ - mred_lazy(p2[n2], F, q, qinv);
    p2[0] = cred(p2[0] << 1, q);
}
/* INTTConjugateInvariant: ring/ntt.go:728-731 */
void lo_intt_ci(const u64 *p1, u64 *p2, int N, u64 ninv, u64 q, u64 qinv, const u64 *roots) {
    intt_ci_core(p1, p2, N, q, qinv, roots);
    for (int i = 0; i < N; i++) p2[i] = mred(p2[i], ninv, q, qinv);
}
/* INTTConjugateInvariantLazy: ring/ntt.go:734-737 */
void lo_intt_ci_lazy(const u64 *p1, u64 *p2, int N, u64 ninv, u64 q, u64 qinv, const u64 *roots) {
    intt_ci_core(p1, p2, N, q, qinv, roots);
    for (int i = 0; i < N; i++) p2[i] = mred_lazy(p2[i], ninv, q, qinv);
}

/* ---- ring/vec_ops.go --------------------------------------------------- */
/* One entry point per reference kernel, selected by opcode. The opcode
 * numbering is shared with include/lattigo_b200.h (LGPU_OP_*). Operand naming
 * follows the reference: p1,p2 inputs, p3 output (read-modify-write for the
 * ThenAdd/ThenSub forms); s0,s1 are the scalar operands. */
enum {
    OP_ADD = 0, OP_ADDLAZY, OP_SUB, OP_SUBLAZY, OP_NEG, OP_REDUCE, OP_REDUCELAZY,
    OP_MULCOEFFSLAZY, OP_MULCOEFFSLAZYTHENADDLAZY,
    OP_MULCOEFFSBARRETT, OP_MULCOEFFSBARRETTLAZY, OP_MULCOEFFSBARRETTTHENADD, OP_MULCOEFFSBARRETTTHENADDLAZY,
    OP_MULCOEFFSMONTGOMERY, OP_MULCOEFFSMONTGOMERYLAZY, OP_MULCOEFFSMONTGOMERYTHENADD,
    OP_MULCOEFFSMONTGOMERYTHENADDLAZY, OP_MULCOEFFSMONTGOMERYLAZYTHENADDLAZY,
    OP_MULCOEFFSMONTGOMERYTHENSUB, OP_MULCOEFFSMONTGOMERYTHENSUBLAZY, OP_MULCOEFFSMONTGOMERYLAZYTHENSUBLAZY,
    OP_MULCOEFFSMONTGOMERYLAZYTHENNEG,
    OP_ADDLAZYTHENMULSCALARMONTGOMERY, OP_ADDSCALARLAZYTHENMULSCALARMONTGOMERY,
    OP_ADDSCALAR, OP_ADDSCALARLAZY, OP_ADDSCALARLAZYTHENNEGTWOMODULUSLAZY, OP_SUBSCALAR,
    OP_MULSCALARMONTGOMERY, OP_MULSCALARMONTGOMERYLAZY, OP_MULSCALARMONTGOMERYTHENADD,
    OP_MULSCALARMONTGOMERYTHENADDSCALAR, OP_SUBTHENMULSCALARMONTGOMERYTWOMODULUS,
    OP_MFORM, OP_MFORMLAZY, OP_IMFORM, OP_ZERO, OP_MASK, OP_COUNT
};

int lo_vecop(int op, const u64 *p1, const u64 *p2, u64 *p3, int N,
             u64 q, u64 qinv, const u64 *brc, u64 s0, u64 s1) {
    u64 twoq = 2 * q;
    switch (op) {
#define LOOP(expr) for (int i = 0; i < N; i++) { u64 x = p1 ? p1[i] : 0, y = p2 ? p2[i] : 0, z = p3[i]; (void)x; (void)y; (void)z; p3[i] = (expr); } break
    case OP_ADD: LOOP(cred(x + y, q));                                   /* vec_ops.go:7  */
    case OP_ADDLAZY: LOOP(x + y);                                        /* :31  */
    case OP_SUB: LOOP(cred((x + q) - y, q));                             /* :55  */
    case OP_SUBLAZY: LOOP(x + q - y);                                    /* :79  */
    case OP_NEG: LOOP(q - x);                                            /* :103 */
    case OP_REDUCE: LOOP(bred_add(x, q, brc));                           /* :125 */
    case OP_REDUCELAZY: LOOP(bred_add_lazy(x, q, brc));                  /* :147 */
    case OP_MULCOEFFSLAZY: LOOP(x * y);                                  /* :169 */
    case OP_MULCOEFFSLAZYTHENADDLAZY: LOOP(z + x * y);                   /* :193 */
    case OP_MULCOEFFSBARRETT: LOOP(bred(x, y, q, brc));                  /* :217 */
    case OP_MULCOEFFSBARRETTLAZY: LOOP(bred_lazy(x, y, q, brc));         /* :241 */
    case OP_MULCOEFFSBARRETTTHENADD: LOOP(cred(z + bred(x, y, q, brc), q)); /* :265 */
    case OP_MULCOEFFSBARRETTTHENADDLAZY: LOOP(z + bred(x, y, q, brc));   /* :289 */
    case OP_MULCOEFFSMONTGOMERY: LOOP(mred(x, y, q, qinv));              /* :313 */
    case OP_MULCOEFFSMONTGOMERYLAZY: LOOP(mred_lazy(x, y, q, qinv));     /* :336 */
    case OP_MULCOEFFSMONTGOMERYTHENADD: LOOP(cred(z + mred(x, y, q, qinv), q)); /* :360 */
    case OP_MULCOEFFSMONTGOMERYTHENADDLAZY: LOOP(z + mred(x, y, q, qinv));      /* :383 */
    case OP_MULCOEFFSMONTGOMERYLAZYTHENADDLAZY: LOOP(z + mred_lazy(x, y, q, qinv)); /* :407 */
    case OP_MULCOEFFSMONTGOMERYTHENSUB: LOOP(cred(z + (q - mred(x, y, q, qinv)), q)); /* :431 */
    case OP_MULCOEFFSMONTGOMERYTHENSUBLAZY: LOOP(z + (q - mred(x, y, q, qinv)));      /* :455 */
    case OP_MULCOEFFSMONTGOMERYLAZYTHENSUBLAZY: LOOP(z + twoq - mred_lazy(x, y, q, qinv)); /* :479 */
    case OP_MULCOEFFSMONTGOMERYLAZYTHENNEG: LOOP(twoq - mred_lazy(x, y, q, qinv));    /* :504 */
    case OP_ADDLAZYTHENMULSCALARMONTGOMERY: LOOP(mred(x + y, s0, q, qinv));           /* :529 (scalarMont=s0) */
    case OP_ADDSCALARLAZYTHENMULSCALARMONTGOMERY: LOOP(mred(x + s0, s1, q, qinv));    /* :553 */
    case OP_ADDSCALAR: LOOP(cred(x + s0, q));                            /* :575 */
    case OP_ADDSCALARLAZY: LOOP(x + s0);                                 /* :597 */
    case OP_ADDSCALARLAZYTHENNEGTWOMODULUSLAZY: LOOP(s0 + twoq - x);     /* :619 */
    case OP_SUBSCALAR: LOOP(cred(x + q - s0, q));                        /* :642 */
    case OP_MULSCALARMONTGOMERY: LOOP(mred(x, s0, q, qinv));             /* :664 */
    case OP_MULSCALARMONTGOMERYLAZY: LOOP(mred_lazy(x, s0, q, qinv));    /* :686 */
    case OP_MULSCALARMONTGOMERYTHENADD: LOOP(cred(z + mred(x, s0, q, qinv), q)); /* :708 */
    case OP_MULSCALARMONTGOMERYTHENADDSCALAR: LOOP(cred(mred(x, s1, q, qinv) + s0, q)); /* :730 */
    case OP_SUBTHENMULSCALARMONTGOMERYTWOMODULUS: LOOP(mred(twoq - y + x, s0, q, qinv)); /* :752 */
    case OP_MFORM: LOOP(mform(x, q, brc));                               /* :778 */
    case OP_MFORMLAZY: LOOP(mform_lazy(x, q, brc));                      /* :800 */
    case OP_IMFORM: LOOP(imform(x, q, qinv));                            /* :822 */
    case OP_ZERO: LOOP(0);                                               /* :847 */
    case OP_MASK: LOOP((x >> s0) & s1);                                  /* :870 (w=s0, mask=s1) */
#undef LOOP
    default: return -1;
    }
    return 0;
}

/* ---- ring/basis_extension.go ------------------------------------------- */

/* ModUpExact: ring/basis_extension.go:282-308 with reconstructRNS (:550-594)
 * and multSum (:597-673).
 *   p1: nQ rows (stride ld1) , p2: nP rows (stride ld2), n coefficients.
 *   Q,mredQ: source moduli;  P,mredP: target moduli.
 *   qoverqiinvqi[nQ]; qoverqimodp[nP][nQ] (row stride ldc); vtimesqmodp[nP][.] (row stride ldv)
 * float64: v = uint64(sum_i float64(y_i)/float64(q_i)), sequential in limb
 * order (:575-593). Built with -ffp-contract=off so no FMA contraction. */
void lo_modup_exact(const u64 *p1, size_t ld1, int nQ, u64 *p2, size_t ld2, int nP, int n,
                    const u64 *Q, const u64 *mredQ, const u64 *P, const u64 *mredP,
                    const u64 *qoverqiinvqi, const u64 *qoverqimodp, size_t ldc,
                    const u64 *vtimesqmodp, size_t ldv) {
    u64 y[64];
    for (int x = 0; x < n; x++) {
        double vi = 0.0;
        for (int i = 0; i < nQ; i++) {
            y[i] = mred(p1[i * ld1 + x], qoverqiinvqi[i], Q[i], mredQ[i]);
            vi += (double)y[i] / (double)Q[i];
        }
        u64 v = (u64)vi;
        for (int j = 0; j < nP; j++) {
            const u64 *c = qoverqimodp + j * ldc;
            u128 acc = (u128)y[0] * c[0];
            u64 rhi = (u64)(acc >> 64), rlo = (u64)acc;
            for (int i = 1; i < nQ; i++) {
                u128 m = (u128)y[i] * c[i];
                u64 mhi = (u64)(m >> 64), mlo = (u64)m;
                u64 s = rlo + mlo; u64 cy = s < rlo; rlo = s;
                rhi += mhi + cy;
            }
            u64 hhi = mulhi(rlo * mredP[j], P[j]);
            p2[j * ld2 + x] = rhi - hhi + P[j] + vtimesqmodp[j * ldv + v];
        }
    }
}

/* Core of Decomposer.DecomposeAndSplit, reconstruction branch:
 * ring/basis_extension.go:438-501 with reconstructRNSCentered (:504-548).
 *   src: the digit's nD source rows (coefficient domain), moduli Qd/mredQd, centred by qhalf[i]
 *   dst rows: nT target rows given by pointer table (rows the reference writes:
 *   all Q limbs outside the digit, then P limbs), with target moduli T/mredT and
 *   the matching rows of qoverqimodp / vtimesqmodp (already selected by caller). */
void lo_decompose_reconstruct(const u64 *src, size_t lds, int nD, u64 *const *dst, int nT, int n,
                              const u64 *Qd, const u64 *mredQd, const u64 *qhalf,
                              const u64 *qoverqiinvqi,
                              const u64 *T, const u64 *mredT,
                              const u64 *const *qoverqimodp_rows, const u64 *const *vtimesqmodp_rows) {
    u64 y[64];
    for (int x = 0; x < n; x++) {
        double vi = 0.0;
        for (int i = 0; i < nD; i++) {
            y[i] = mred(src[i * lds + x] + qhalf[i], qoverqiinvqi[i], Qd[i], mredQd[i]);
            vi += (double)y[i] / (double)Qd[i];
        }
        u64 v = (u64)vi;
        for (int j = 0; j < nT; j++) {
            const u64 *c = qoverqimodp_rows[j];
            u128 acc = (u128)y[0] * c[0];
            u64 rhi = (u64)(acc >> 64), rlo = (u64)acc;
            for (int i = 1; i < nD; i++) {
                u128 m = (u128)y[i] * c[i];
                u64 mhi = (u64)(m >> 64), mlo = (u64)m;
                u64 s = rlo + mlo; u64 cy = s < rlo; rlo = s;
                rhi += mhi + cy;
            }
            u64 hhi = mulhi(rlo * mredT[j], T[j]);
            dst[j][x] = rhi - hhi + T[j] + vtimesqmodp_rows[j][v];
        }
    }
}

/* Single-limb digit fast path of DecomposeAndSplit: ring/basis_extension.go:402-436 */
void lo_decompose_single(const u64 *src, u64 qsrc, u64 *const *dst, int nT, int n,
                         const u64 *T, const u64 *brcT /* 2 per target */) {
    for (int j = 0; j < n; j++) {
        u64 coeff = src[j], pos = 1, neg = 0;
        if (coeff >= (qsrc >> 1)) { coeff = qsrc - coeff; pos = 0; neg = 1; }
        for (int i = 0; i < nT; i++) {
            u64 tmp = bred_add(coeff, T[i], brcT + 2 * i);
            dst[i][j] = tmp * pos + (T[i] - tmp) * neg;
        }
    }
}

/* ---- ring/automorphism.go ---------------------------------------------- */
static inline u64 bitrev64(u64 x, int bits) {
    u64 r = 0;
    for (int i = 0; i < 64; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return bits ? r >> (64 - bits) : 0;
}
/* AutomorphismNTTIndex: ring/automorphism.go:12-34 */
void lo_automorphism_ntt_index(int N, u64 nthroot, u64 galel, u64 *index) {
    int lg = 0; while ((1ull << (lg + 1)) < nthroot) lg++;   /* bits.Len64(NthRoot-1)-1 */
    u64 mask = nthroot - 1;
    for (int i = 0; i < N; i++) {
        u64 tmp1 = 2 * bitrev64((u64)i, lg) + 1;
        u64 tmp2 = ((galel * tmp1 & mask) - 1) >> 1;
        index[i] = bitrev64(tmp2, lg);
    }
}
/* AutomorphismNTTWithIndex[ThenAddLazy]: ring/automorphism.go:50-109 (one row) */
void lo_automorphism_ntt_row(const u64 *in, const u64 *index, u64 *out, int N, int accumulate) {
    if (accumulate) for (int j = 0; j < N; j++) out[j] += in[index[j]];
    else            for (int j = 0; j < N; j++) out[j] = in[index[j]];
}
/* Ring.Automorphism (coefficient domain, ConjugateInvariant ring): ring/automorphism.go:123-155 */
void lo_automorphism_row_ci(const u64 *in, u64 *out, int N, u64 gen, u64 q) {
    u64 n = (u64)N, mask = 2 * n - 1; int logN = 0; while (((u64)1 << logN) <= mask) logN++;   /* bits.Len64(mask) */
    for (u64 i = 0; i < 2 * n; i++) {
        u64 raw = i * gen, index = raw & mask, tmp = (raw >> logN) & 1;
        if (index < n) {
            u64 idx = i;
            if (idx >= n) { idx = 2 * n - idx; tmp ^= 1; }
            out[index] = in[idx] * (tmp ^ 1) | (q - in[idx]) * tmp;
        }
    }
}
/* Ring.Automorphism (coefficient domain, Standard ring): ring/automorphism.go:158-175 */
void lo_automorphism_row(const u64 *in, u64 *out, int N, u64 gen, u64 q) {
    u64 mask = (u64)N - 1; int logN = 0; while ((1 << logN) < N) logN++;
    for (u64 i = 0; i < (u64)N; i++) {
        u64 raw = i * gen, idx = raw & mask, tmp = (raw >> logN) & 1;
        out[idx] = in[i] * (tmp ^ 1) | (q - in[i]) * tmp;
    }
}
