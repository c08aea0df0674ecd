/*
 * oracle/lattigo_cpu_batch.c -- TEST / BASELINE INFRASTRUCTURE ONLY (never linked into lattigo_b200/).
 *
 * The CPU arm of bench.py ("restated reference (C), N cores"): one native batch loop over ciphertext pairs for the
 * measured op sequence  ckks.Evaluator.MulRelinNew + Rescale  (schemes/ckks/evaluator.go:719-872, :477-515) with the
 * multiple-P key-switch (core/rlwe/evaluator_gadget_product.go:129-201, :39-97, :487-510). Go is not installed in this
 * image, so the reference cannot run; this file mirrors how the reference runs the workload:
 *   - b.RunParallel: one goroutine per ciphertext, all goroutines share the read-only inputs and keys
 *     (schemes/ckks/ckks_benchmarks_test.go:218-229)  ->  one pthread per pair slot, shared inputs / key;
 *   - sync.Pool scratch recycling (ring/pool.go:10-61)   ->  one workspace per thread, allocated once and reused
 *     for every pair and every call (first-touched by its own thread).
 * The arithmetic is the oracle's (lattigo_oracle.c is #included, same scalar primitives and the same lazy schedules), so
 * tests/test_cpu_batch.py can require bit-equality with oracle.py's CKKSEvaluator. Only the loop nests of the
 * transforms are specialised for short inner strides (the reference hand-unrolls them: ring/ntt.go:258-552, :608-714);
 * every butterfly and every reduction is the same operation in the same stage order.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdlib.h>
#include <time.h>
#include "lattigo_oracle.c"

typedef struct {
    int N, nQ, nP, nd;              /* limbs in use: Q[0..nQ), P[0..nP); nd = BaseRNSDecompositionVectorSize */
    int qi_overf, pi_overf;         /* QiOverflowMargin(level)>>1, PiOverflowMargin(levelP)>>1 */
    const u64 *mod, *qinv, *bred, *ninv;       /* [nQ+nP] (bred: 2 per limb), Q limbs first */
    const u64 *const *roots_fwd;               /* [nQ+nP] */
    const u64 *const *roots_bwd;
    const u64 *evk;                            /* [nd][2][nQ+nP][N], NTT + Montgomery */
    /* Decomposer (ring/basis_extension.go:381-502), per digit d */
    const int *dig_start, *dig_n;              /* [nd] first source limb, number of source limbs (1 => single-limb rule) */
    int dmax;                                  /* row length of the per-digit tables below (>= max dig_n) */
    const u64 *dec_qhalf;                      /* [nd][dmax]  floor(Qd/2) mod q_i, source limbs */
    const u64 *dec_inv;                        /* [nd][dmax]  qoverqiinvqi */
    const u64 *dec_c;                          /* [nd][nQ+nP][dmax]   qoverqimodp row of target limb j */
    const u64 *dec_v;                          /* [nd][nQ+nP][dmax+1] vtimesqmodp row of target limb j */
    const u64 *dec_half_t;                     /* [nd][nQ+nP] floor(Qd/2) mod t_j */
    /* ModDownQPtoQNTT (ring/basis_extension.go:235-256) at (levelQ, levelP) */
    const u64 *md_phalf_p, *md_inv;            /* [nP] */
    const u64 *md_c;                           /* [nQ][nP] */
    const u64 *md_v;                           /* [nQ][nP+1] */
    const u64 *md_phalf_q;                     /* [nQ] */
    const u64 *md_scal;                        /* [nQ] q_i - modDownConstantsPtoQ[levelP][i] */
    /* DivRoundByLastModulusNTT (ring/scaling.go:101-122) */
    const u64 *rescale;                        /* [nQ-1] RescaleConstants[level-1][i] */
} lo_plan;

/* ---- transforms: same butterflies / stage order as lo_ntt_lazy / intt_core, loop nests specialised ------------ */
static void ntt_lazy_fast(const u64 *p1, u64 *p2, int N, u64 q, u64 qinv, const u64 *restrict roots) {
    int t = N >> 1;
    u64 F = roots[1];
    for (int j = 0; j < t; j++) {                   /* stage m = 1: no U correction (ring/ntt.go:275-310) */
        u64 U = p1[j], V = mred_lazy(p1[j + t], F, q, qinv);
        p2[j] = U + V; p2[j + t] = U + 2 * q - V;
    }
    for (int m = 2; m < N; m <<= 1) {
        t >>= 1;
        int len = 0; for (int mm = m; mm; mm >>= 1) len++;
        const int reduce = (t == 1) || (len & 1);    /* ring/ntt.go:318, :500-518 */
        if (t >= 4) {
            for (int i = 0; i < m; i++) {
                u64 *x = p2 + ((size_t)(i * t) << 1), *y = x + t;
                F = roots[m + i];
                if (reduce) for (int j = 0; j < t; j++) bfly(&x[j], &y[j], F, q, qinv, 1);
                else        for (int j = 0; j < t; j++) bfly(&x[j], &y[j], F, q, qinv, 0);
            }
        } else if (t == 2) {
            for (int i = 0; i < m; i++) {
                u64 *x = p2 + 4 * (size_t)i; F = roots[m + i];
                bfly(&x[0], &x[2], F, q, qinv, reduce); bfly(&x[1], &x[3], F, q, qinv, reduce);
            }
        } else {
            for (int i = 0; i < m; i++) { u64 *x = p2 + 2 * (size_t)i; bfly(&x[0], &x[1], roots[m + i], q, qinv, 1); }
        }
    }
}
static void ntt_fast(const u64 *p1, u64 *p2, int N, u64 q, u64 qinv, const u64 *brc, const u64 *roots) {
    ntt_lazy_fast(p1, p2, N, q, qinv, roots);
    for (int i = 0; i < N; i++) p2[i] = bred_add(p2[i], q, brc);      /* reducevec, ring/ntt.go:176 */
}
/* INTTStandard / INTTStandardLazy for N >= 16 (identical: ring/ntt.go:185-206) */
static void intt_fast(const u64 *p1, u64 *p2, int N, u64 ninv, u64 q, u64 qinv, const u64 *restrict roots) {
    int h = N >> 1;
    for (int i = 0; i < h; i++) {                  /* t = 1 */
        u64 x = p1[2 * i], y = p1[2 * i + 1];
        ibfly(&x, &y, roots[h + i], q, qinv);
        p2[2 * i] = x; p2[2 * i + 1] = y;
    }
    int t = 2;
    for (int m = N >> 1; m > 1; m >>= 1) {
        h = m >> 1;
        if (t == 2) {
            for (int i = 0; i < h; i++) {
                u64 *x = p2 + 4 * (size_t)i; const u64 F = roots[h + i];
                ibfly(&x[0], &x[2], F, q, qinv); ibfly(&x[1], &x[3], F, q, qinv);
            }
        } else {
            for (int i = 0; i < h; i++) {
                u64 *x = p2 + (size_t)i * 2 * t, *y = x + t; const u64 F = roots[h + i];
                for (int j = 0; j < t; j++) ibfly(&x[j], &y[j], F, q, qinv);
            }
        }
        t <<= 1;
    }
    for (int i = 0; i < N; i++) p2[i] = mred(p2[i], ninv, q, qinv);   /* mulscalarmontgomeryvec, ring/vec_ops.go:664 */
}

/* exported single-row entry points so the tests can pin the specialised loop nests on the golden vectors */
void lb_ntt(const u64 *p1, u64 *p2, int N, u64 q, u64 qinv, const u64 *brc, const u64 *roots) { ntt_fast(p1, p2, N, q, qinv, brc, roots); }
void lb_ntt_lazy(const u64 *p1, u64 *p2, int N, u64 q, u64 qinv, const u64 *roots) { ntt_lazy_fast(p1, p2, N, q, qinv, roots); }
void lb_intt(const u64 *p1, u64 *p2, int N, u64 ninv, u64 q, u64 qinv, const u64 *roots) { intt_fast(p1, p2, N, ninv, q, qinv, roots); }

/* ---- per-thread workspace (sync.Pool analogue) ------------------------------------------------------------------- */
typedef struct {
    size_t words;
    u64 *mem;
} lo_ws;
#define LO_MAX_THREADS 1024
static lo_ws g_ws[LO_MAX_THREADS];

static u64 *ws_get(int tid, size_t rows, int N) {
    lo_ws *w = &g_ws[tid];
    const size_t words = rows * (size_t)N;
    if (w->words < words) {
        free(w->mem);
        if (posix_memalign((void **)&w->mem, 64, words * sizeof(u64))) { w->mem = NULL; w->words = 0; return NULL; }
        w->words = words;
        memset(w->mem, 0, words * sizeof(u64));                  /* first touch on the owning thread */
    }
    return w->mem;
}
void lo_batch_release(void) {
    for (int i = 0; i < LO_MAX_THREADS; i++) { free(g_ws[i].mem); g_ws[i].mem = NULL; g_ws[i].words = 0; }
}

/* ---- one ciphertext pair ------------------------------------------------------------------------------------------ */
#define ROW(base, i) ((base) + (size_t)(i) * N)
#define FWD(g) (P->roots_fwd[g])
#define BWD(g) (P->roots_bwd[g])

/* DecomposeSingleNTT (core/rlwe/evaluator_gadget_product.go:487-510) for digit d: cQ[nQ], cP[nP] <- NTT rows */
static void decompose_single_ntt(const lo_plan *P, int d, const u64 *cxNTT, const u64 *cxInv, u64 *cQ, u64 *cP) {
    const int N = P->N, nQ = P->nQ, nP = P->nP, nT = nQ + nP;
    const int st = P->dig_start[d], nD = P->dig_n[d], dmax = P->dmax;
    const u64 *mod = P->mod, *qinv = P->qinv;
    if (nD == 1) {
        /* single-limb digit: ring/basis_extension.go:402-436 (centre, reduce into every Q and P limb) */
        const u64 qs = mod[st];
        const u64 *src = ROW(cxInv, st);
        for (int j = 0; j < N; j++) {
            u64 coeff = src[j], pos = 1, neg = 0;
            if (coeff >= (qs >> 1)) { coeff = qs - coeff; pos = 0; neg = 1; }
            for (int g = 0; g < nT; g++) {
                u64 *dst = g < nQ ? ROW(cQ, g) : ROW(cP, g - nQ);
                const u64 tmp = bred_add(coeff, mod[g], P->bred + 2 * g);
                dst[j] = tmp * pos + (mod[g] - tmp) * neg;
            }
        }
    } else {
        /* reconstruction branch :438-501 with reconstructRNSCentered :504-548 */
        const u64 *qh = P->dec_qhalf + (size_t)d * dmax, *inv = P->dec_inv + (size_t)d * dmax;
        const u64 *C = P->dec_c + (size_t)d * nT * dmax, *V = P->dec_v + (size_t)d * nT * (dmax + 1);
        const u64 *ht = P->dec_half_t + (size_t)d * nT;
        u64 y[64];
        for (int x = 0; x < N; x++) {
            double vi = 0.0;
            for (int i = 0; i < nD; i++) {
                y[i] = mred(ROW(cxInv, st + i)[x] + qh[i], inv[i], mod[st + i], qinv[st + i]);
                vi += (double)y[i] / (double)mod[st + i];
            }
            const u64 v = (u64)vi;
            for (int g = 0; g < nT; g++) {
                if (g >= st && g < st + nD) continue;             /* the digit's own rows are copied below */
                const u64 *c = C + (size_t)g * dmax;
                u128 acc = (u128)y[0] * c[0];
                u64 rhi = (u64)(acc >> 64), rlo = (u64)acc;
                for (int i = 1; i < nD; i++) {
                    u128 m = (u128)y[i] * c[i];
                    u64 mhi = (u64)(m >> 64), mlo = (u64)m;
                    u64 s = rlo + mlo; u64 cy = s < rlo; rlo = s;
                    rhi += mhi + cy;
                }
                const u64 t = mod[g];
                u64 r = rhi - mulhi(rlo * qinv[g], t) + t + V[(size_t)g * (dmax + 1) + v];
                r = cred(r + t - ht[g], t);                        /* SubScalarBigint(QHalf), :499-500 */
                (g < nQ ? ROW(cQ, g) : ROW(cP, g - nQ))[x] = r;
            }
        }
    }
    /* :494-508: own rows are copied from the NTT-domain input, every other row is transformed */
    for (int g = 0; g < nQ; g++) {
        if (g >= st && g < st + nD) {
            /* p0idxst <= x < p0idxst + nbPi, and only rows <= levelQ exist: exactly the digit's nD source limbs */
            memcpy(ROW(cQ, g), ROW(cxNTT, g), (size_t)N * sizeof(u64));
        } else {
            ntt_fast(ROW(cQ, g), ROW(cQ, g), N, mod[g], qinv[g], P->bred + 2 * g, FWD(g));
        }
    }
    for (int j = 0; j < nP; j++) {
        const int g = nQ + j;
        ntt_fast(ROW(cP, j), ROW(cP, j), N, mod[g], qinv[g], P->bred + 2 * g, FWD(g));
    }
}

static void reduce_rows(const lo_plan *P, u64 *a, int g0, int n) {
    const int N = P->N;
    for (int i = 0; i < n; i++) {
        u64 *r = ROW(a, i); const u64 q = P->mod[g0 + i]; const u64 *b = P->bred + 2 * (g0 + i);
        for (int x = 0; x < N; x++) r[x] = bred_add(r[x], q, b);
    }
}

/* ModDownQPtoQNTT for one component: out[i] = (accQ[i] - ModUpPtoQ(INTT(accP))[i]) * P^-1, ring/basis_extension.go:235-256 */
static void moddown_ntt(const lo_plan *P, const u64 *accQ, const u64 *accP, u64 *buffP, u64 *buffQ, u64 *out) {
    const int N = P->N, nQ = P->nQ, nP = P->nP;
    const u64 *mod = P->mod, *qinv = P->qinv;
    for (int j = 0; j < nP; j++) {
        const int g = nQ + j;
        intt_fast(ROW(accP, j), ROW(buffP, j), N, P->ninv[g], mod[g], qinv[g], BWD(g));        /* INTTLazy */
        u64 *r = ROW(buffP, j); const u64 s = P->md_phalf_p[j], q = mod[g];
        for (int x = 0; x < N; x++) r[x] = cred(r[x] + s, q);                                   /* AddScalarBigint(PHalf) */
    }
    u64 y[64];
    for (int x = 0; x < N; x++) {                                                               /* ModUpExact :282-308 */
        double vi = 0.0;
        for (int j = 0; j < nP; j++) {
            y[j] = mred(ROW(buffP, j)[x], P->md_inv[j], mod[nQ + j], qinv[nQ + j]);
            vi += (double)y[j] / (double)mod[nQ + j];
        }
        const u64 v = (u64)vi;
        for (int i = 0; i < nQ; i++) {
            const u64 *c = P->md_c + (size_t)i * nP;
            u128 acc = (u128)y[0] * c[0];
            u64 rhi = (u64)(acc >> 64), rlo = (u64)acc;
            for (int j = 1; j < nP; j++) {
                u128 m = (u128)y[j] * c[j];
                u64 mhi = (u64)(m >> 64), mlo = (u64)m;
                u64 s = rlo + mlo; u64 cy = s < rlo; rlo = s;
                rhi += mhi + cy;
            }
            const u64 q = mod[i];
            u64 r = rhi - mulhi(rlo * qinv[i], q) + q + P->md_v[(size_t)i * (nP + 1) + v];
            ROW(buffQ, i)[x] = cred(r + q - P->md_phalf_q[i], q);                               /* SubScalarBigint(PHalf) */
        }
    }
    for (int i = 0; i < nQ; i++) {
        const u64 q = mod[i], qi = qinv[i], s0 = P->md_scal[i], twoq = 2 * q;
        u64 *e = ROW(buffQ, i);
        ntt_lazy_fast(e, e, N, q, qi, FWD(i));                                                  /* NTTLazy */
        const u64 *a = ROW(accQ, i); u64 *o = ROW(out, i);
        for (int x = 0; x < N; x++) o[x] = mred(twoq - a[x] + e[x], s0, q, qi);                 /* SubThenMulScalarMontgomeryTwoModulus */
    }
}

/* DivRoundByLastModulusNTT, ring/scaling.go:101-122 : p0[nQ rows] -> p1[nQ-1 rows] */
static void div_round_last_ntt(const lo_plan *P, const u64 *p0, u64 *p1, u64 *buff0, u64 *buff1) {
    const int N = P->N, L = P->nQ - 1;
    const u64 *mod = P->mod, *qinv = P->qinv;
    intt_fast(ROW(p0, L), buff0, N, P->ninv[L], mod[L], qinv[L], BWD(L));
    const u64 pHalf = (mod[L] - 1) >> 1;
    for (int x = 0; x < N; x++) buff0[x] = cred(buff0[x] + pHalf, mod[L]);
    for (int i = 0; i < L; i++) {
        const u64 q = mod[i], qi = qinv[i], s = q - pHalf % q, rc = P->rescale[i], twoq = 2 * q;
        for (int x = 0; x < N; x++) buff1[x] = buff0[x] + s;                                     /* AddScalarLazy */
        ntt_lazy_fast(buff1, buff1, N, q, qi, FWD(i));
        const u64 *a = ROW(p0, i); u64 *o = ROW(p1, i);
        for (int x = 0; x < N; x++) o[x] = mred(twoq - a[x] + buff1[x], rc, q, qi);
    }
}

/* rows of workspace one pair needs */
static size_t ws_rows(const lo_plan *P) { return (size_t)9 * P->nQ + (size_t)4 * P->nP + 2; }

/* a0,a1,b0,b1: [nQ][N]; out0,out1: [nQ-1][N] (rescaled) ; mid0/mid1 (optional, [nQ][N]): the MulRelinNew result */
static void one_pair(const lo_plan *P, u64 *ws, const u64 *a0, const u64 *a1, const u64 *b0, const u64 *b1,
                     u64 *out0, u64 *out1, u64 *mid0, u64 *mid1) {
    const int N = P->N, nQ = P->nQ, nP = P->nP, nT = nQ + nP;
    const u64 *mod = P->mod, *qinv = P->qinv;
    u64 *c0 = ws, *c1 = c0 + (size_t)nQ * N, *c2 = c1 + (size_t)nQ * N, *cxInv = c2 + (size_t)nQ * N;
    u64 *cQ = cxInv + (size_t)nQ * N, *accQ0 = cQ + (size_t)nQ * N, *accQ1 = accQ0 + (size_t)nQ * N;
    u64 *buffQ = accQ1 + (size_t)nQ * N, *c00 = buffQ + (size_t)nQ * N;
    u64 *cP = c00 + (size_t)nQ * N, *accP0 = cP + (size_t)nP * N, *accP1 = accP0 + (size_t)nP * N, *buffP = accP1 + (size_t)nP * N;
    u64 *b0r = buffP + (size_t)nP * N, *b1r = b0r + N;
    u64 *c01 = cQ;    /* cQ is free until the key-switch starts */
    /* tensor, schemes/ckks/evaluator.go:807-820 */
    for (int i = 0; i < nQ; i++) {
        const u64 q = mod[i], qi = qinv[i]; const u64 *br = P->bred + 2 * i;
        const u64 *x0 = ROW(a0, i), *x1 = ROW(a1, i), *y0 = ROW(b0, i), *y1 = ROW(b1, i);
        u64 *m0 = ROW(c00, i), *m1 = ROW(c01, i), *r0 = ROW(c0, i), *r1 = ROW(c1, i), *r2 = ROW(c2, i);
        for (int x = 0; x < N; x++) m0[x] = mform(x0[x], q, br);
        for (int x = 0; x < N; x++) m1[x] = mform(x1[x], q, br);
        for (int x = 0; x < N; x++) r0[x] = mred(m0[x], y0[x], q, qi);
        for (int x = 0; x < N; x++) r2[x] = mred(m1[x], y1[x], q, qi);
        for (int x = 0; x < N; x++) r1[x] = mred(m0[x], y1[x], q, qi);
        for (int x = 0; x < N; x++) r1[x] = cred(r1[x] + mred(m1[x], y0[x], q, qi), q);
    }
    /* GadgetProduct(level, c2, rlk): gadgetProductMultiplePLazy, evaluator_gadget_product.go:129-201 */
    for (int i = 0; i < nQ; i++) intt_fast(ROW(c2, i), ROW(cxInv, i), N, P->ninv[i], mod[i], qinv[i], BWD(i));
    int reduce = 0;
    for (int d = 0; d < P->nd; d++) {
        decompose_single_ntt(P, d, c2, cxInv, cQ, cP);
        for (int comp = 0; comp < 2; comp++) {
            const u64 *k = P->evk + ((size_t)d * 2 + comp) * nT * N;
            u64 *aQ = comp ? accQ1 : accQ0, *aP = comp ? accP1 : accP0;
            for (int g = 0; g < nT; g++) {
                const u64 q = mod[g], qi = qinv[g];
                const u64 *kr = ROW(k, g), *cr = g < nQ ? ROW(cQ, g) : ROW(cP, g - nQ);
                u64 *ar = g < nQ ? ROW(aQ, g) : ROW(aP, g - nQ);
                if (d == 0) for (int x = 0; x < N; x++) ar[x] = mred_lazy(kr[x], cr[x], q, qi);           /* MulCoeffsMontgomeryLazy */
                else        for (int x = 0; x < N; x++) ar[x] = ar[x] + mred_lazy(kr[x], cr[x], q, qi);   /* ...LazyThenAddLazy */
            }
        }
        if (reduce % P->qi_overf == P->qi_overf - 1) { reduce_rows(P, accQ0, 0, nQ); reduce_rows(P, accQ1, 0, nQ); }
        if (reduce % P->pi_overf == P->pi_overf - 1) { reduce_rows(P, accP0, nQ, nP); reduce_rows(P, accP1, nQ, nP); }
        reduce++;
    }
    if (reduce % P->qi_overf != 0) { reduce_rows(P, accQ0, 0, nQ); reduce_rows(P, accQ1, 0, nQ); }
    if (reduce % P->pi_overf != 0) { reduce_rows(P, accP0, nQ, nP); reduce_rows(P, accP1, nQ, nP); }
    /* ModDown :39-97 (NTT in, NTT out), then ringQ.Add x2 (schemes/ckks/evaluator.go:837-838) */
    u64 *t0 = cxInv, *t1 = cQ;
    moddown_ntt(P, accQ0, accP0, buffP, buffQ, t0);
    moddown_ntt(P, accQ1, accP1, buffP, buffQ, t1);
    for (int i = 0; i < nQ; i++) {
        const u64 q = mod[i];
        u64 *r0 = ROW(c0, i), *r1 = ROW(c1, i); const u64 *s0 = ROW(t0, i), *s1 = ROW(t1, i);
        for (int x = 0; x < N; x++) r0[x] = cred(r0[x] + s0[x], q);
        for (int x = 0; x < N; x++) r1[x] = cred(r1[x] + s1[x], q);
    }
    if (mid0) memcpy(mid0, c0, (size_t)nQ * N * sizeof(u64));
    if (mid1) memcpy(mid1, c1, (size_t)nQ * N * sizeof(u64));
    /* Rescale :477-515 */
    u64 *o0 = out0 ? out0 : accQ0, *o1 = out1 ? out1 : accQ1;
    div_round_last_ntt(P, c0, o0, b0r, b1r);
    div_round_last_ntt(P, c1, o1, b0r, b1r);
}

/* ---- batch driver -------------------------------------------------------------------------------------------------- */
typedef struct {
    const lo_plan *P;
    const u64 *a, *b; u64 *out, *mid;
    size_t in_stride, out_stride, mid_stride;     /* words between pairs (0 = every pair reads the same inputs) */
    int npairs, tid;
    volatile int *next;
    double *pair_seconds;
    int pin_cpu;
} lo_job;

static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

static void *worker(void *arg) {
    lo_job *J = (lo_job *)arg;
    const lo_plan *P = J->P;
    if (J->pin_cpu >= 0) {
        cpu_set_t set; CPU_ZERO(&set); CPU_SET(J->pin_cpu, &set);
        pthread_setaffinity_np(pthread_self(), sizeof(set), &set);
    }
    u64 *ws = ws_get(J->tid, ws_rows(P), P->N);
    if (!ws) return (void *)1;
    const size_t poly = (size_t)P->nQ * P->N, opoly = (size_t)(P->nQ - 1) * P->N;
    for (;;) {
        const int i = __sync_fetch_and_add(J->next, 1);
        if (i >= J->npairs) break;
        const double t0 = now_s();
        const u64 *a = J->a + (size_t)i * J->in_stride, *b = J->b + (size_t)i * J->in_stride;
        u64 *o = J->out ? J->out + (size_t)i * J->out_stride : NULL;
        u64 *m = J->mid ? J->mid + (size_t)i * J->mid_stride : NULL;
        one_pair(P, ws, a, a + poly, b, b + poly, o, o ? o + opoly : NULL, m, m ? m + poly : NULL);
        if (J->pair_seconds) J->pair_seconds[i] = now_s() - t0;
    }
    return NULL;
}

/* a, b: [npairs or 1][2][nQ][N]; out: [npairs][2][nQ-1][N] or NULL (results stay in the thread's workspace);
 * mid: optional [npairs][2][nQ][N] MulRelinNew result. cpus: optional list of CPU ids to pin the threads to
 * (nthreads entries) or NULL. Returns wall seconds, < 0 on error. */
double lo_ckks_mulrelin_rescale_batch(const lo_plan *P, const u64 *a, const u64 *b, size_t in_stride, u64 *out, u64 *mid,
                                      int npairs, int nthreads, const int *cpus, double *pair_seconds) {
    if (nthreads < 1 || nthreads > LO_MAX_THREADS || P->nP < 2 || P->nQ < 2 || P->N < 16) return -1.0;
    volatile int next = 0;
    pthread_t th[LO_MAX_THREADS];
    static lo_job jobs[LO_MAX_THREADS];
    const size_t poly = (size_t)P->nQ * P->N, opoly = (size_t)(P->nQ - 1) * P->N;
    const double t0 = now_s();
    for (int t = 0; t < nthreads; t++) {
        lo_job j = {P, a, b, out, mid, in_stride, 2 * opoly, 2 * poly, npairs, t, &next, pair_seconds, cpus ? cpus[t] : -1};
        jobs[t] = j;
        if (pthread_create(&th[t], NULL, worker, &jobs[t])) return -2.0;
    }
    int bad = 0;
    for (int t = 0; t < nthreads; t++) { void *r; pthread_join(th[t], &r); bad |= (r != NULL); }
    return bad ? -3.0 : now_s() - t0;
}
