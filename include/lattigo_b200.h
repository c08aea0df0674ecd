/*
 * lattigo_b200.h -- C ABI of the B200-native RNS polynomial-ring engine.
 *
 * Drop-in boundary for ONE hot path of tuneinsight/lattigo v6.2.0: the `ring` package's NTT /
 * coefficient-wise / basis-extension / rescale / automorphism loops and core/rlwe.Evaluator's
 * GadgetProduct key-switch family. The reference has no FFI of its own (pure Go); each entry point
 * below names the Go method it replaces (file:line relative to the upstream tree) and is what a thin
 * cgo shim in forked `ring`, `ring/ringqp` and `core/rlwe` packages binds.
 *
 * Conventions
 *   - Polynomials are `uint64_t*` DEVICE-ACCESSIBLE pointers (cudaMalloc / managed / lgpu_malloc) to a
 *     contiguous (limb, coeff) row-major block: row i = residues mod the i-th modulus, N words per row.
 *     ring.Poly{Coeffs [][]uint64} (ring/poly.go:13-15) maps onto it with Coeffs[i] = row i.
 *   - `ring` selects the moduli chain: LGPU_RING_Q or LGPU_RING_P (ringqp.Ring{RingQ,RingP},
 *     ring/ringqp/ring.go:15-17). `level` is the reference's level (number of limbs - 1).
 *   - All calls are asynchronous on `stream` (a cudaStream_t cast to void*; NULL = the CUDA default
 *     stream); lgpu_sync() before host access. Entry points ending in _host take HOST pointers and
 *     include the copies.
 *   - Return 0 on success, non-zero on error with lgpu_last_error() (thread-local) describing it.
 *     The shim maps ring-level errors to panic and evaluator-level errors to `error` like the
 *     reference (ring/ntt.go:212, core/rlwe/evaluator_gadget_product.go:110-112).
 *   - Outputs are bit-identical to the reference on identical inputs: canonical residues for the
 *     non-Lazy methods, the reference's exact representative for the *Lazy ones.
 */
#ifndef LATTIGO_B200_H
#define LATTIGO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lgpu_ctx lgpu_ctx;

#define LGPU_RING_Q 0
#define LGPU_RING_P 1

#define LGPU_RING_STANDARD 0           /* ring.Standard            (ring/ring.go:27-31) */
#define LGPU_RING_CONJUGATE_INVARIANT 1 /* ring.ConjugateInvariant  */

/* Opcodes of lgpu_vecop / lgpu_subring_vecop: one per kernel of ring/vec_ops.go, named after the
 * SubRing method that wraps it (ring/subring_ops.go:6-273). */
enum lgpu_vecop_code {
    LGPU_OP_ADD = 0,                             /* SubRing.Add                              vec_ops.go:7   */
    LGPU_OP_ADDLAZY,                             /* AddLazy                                  :31  */
    LGPU_OP_SUB,                                 /* Sub                                      :55  */
    LGPU_OP_SUBLAZY,                             /* SubLazy                                  :79  */
    LGPU_OP_NEG,                                 /* Neg                                      :103 */
    LGPU_OP_REDUCE,                              /* Reduce                                   :125 */
    LGPU_OP_REDUCELAZY,                          /* ReduceLazy                               :147 */
    LGPU_OP_MULCOEFFSLAZY,                       /* MulCoeffsLazy                            :169 */
    LGPU_OP_MULCOEFFSLAZYTHENADDLAZY,            /* MulCoeffsLazyThenAddLazy                 :193 */
    LGPU_OP_MULCOEFFSBARRETT,                    /* MulCoeffsBarrett                         :217 */
    LGPU_OP_MULCOEFFSBARRETTLAZY,                /* MulCoeffsBarrettLazy                     :241 */
    LGPU_OP_MULCOEFFSBARRETTTHENADD,             /* MulCoeffsBarrettThenAdd                  :265 */
    LGPU_OP_MULCOEFFSBARRETTTHENADDLAZY,         /* MulCoeffsBarrettThenAddLazy              :289 */
    LGPU_OP_MULCOEFFSMONTGOMERY,                 /* MulCoeffsMontgomery                      :313 */
    LGPU_OP_MULCOEFFSMONTGOMERYLAZY,             /* MulCoeffsMontgomeryLazy                  :336 */
    LGPU_OP_MULCOEFFSMONTGOMERYTHENADD,          /* MulCoeffsMontgomeryThenAdd               :360 */
    LGPU_OP_MULCOEFFSMONTGOMERYTHENADDLAZY,      /* MulCoeffsMontgomeryThenAddLazy           :383 */
    LGPU_OP_MULCOEFFSMONTGOMERYLAZYTHENADDLAZY,  /* MulCoeffsMontgomeryLazyThenAddLazy       :407 */
    LGPU_OP_MULCOEFFSMONTGOMERYTHENSUB,          /* MulCoeffsMontgomeryThenSub               :431 */
    LGPU_OP_MULCOEFFSMONTGOMERYTHENSUBLAZY,      /* MulCoeffsMontgomeryThenSubLazy           :455 */
    LGPU_OP_MULCOEFFSMONTGOMERYLAZYTHENSUBLAZY,  /* MulCoeffsMontgomeryLazyThenSubLazy       :479 */
    LGPU_OP_MULCOEFFSMONTGOMERYLAZYTHENNEG,      /* MulCoeffsMontgomeryLazyThenNeg           :504 */
    LGPU_OP_ADDLAZYTHENMULSCALARMONTGOMERY,      /* AddLazyThenMulScalarMontgomery (s0)      :529 */
    LGPU_OP_ADDSCALARLAZYTHENMULSCALARMONTGOMERY,/* AddScalarLazyThenMulScalarMontgomery(s0,s1) :553 */
    LGPU_OP_ADDSCALAR,                           /* AddScalar (s0)                           :575 */
    LGPU_OP_ADDSCALARLAZY,                       /* AddScalarLazy (s0)                       :597 */
    LGPU_OP_ADDSCALARLAZYTHENNEGTWOMODULUSLAZY,  /* AddScalarLazyThenNegTwoModulusLazy (s0)  :619 */
    LGPU_OP_SUBSCALAR,                           /* SubScalar (s0)                           :642 */
    LGPU_OP_MULSCALARMONTGOMERY,                 /* MulScalarMontgomery (s0)                 :664 */
    LGPU_OP_MULSCALARMONTGOMERYLAZY,             /* MulScalarMontgomeryLazy (s0)             :686 */
    LGPU_OP_MULSCALARMONTGOMERYTHENADD,          /* MulScalarMontgomeryThenAdd (s0)          :708 */
    LGPU_OP_MULSCALARMONTGOMERYTHENADDSCALAR,    /* MulScalarMontgomeryThenAddScalar (s0,s1=scalarMont) :730 */
    LGPU_OP_SUBTHENMULSCALARMONTGOMERYTWOMODULUS,/* SubThenMulScalarMontgomeryTwoModulus (s0): (p1 - p2 + 2q) * s0  :752 */
    LGPU_OP_MFORM,                               /* MForm                                    :778 */
    LGPU_OP_MFORMLAZY,                           /* MFormLazy                                :800 */
    LGPU_OP_IMFORM,                              /* IMForm                                   :822 */
    LGPU_OP_ZERO,                                /* ZeroVec                                  :847 */
    LGPU_OP_MASK,                                /* MaskVec (w = s0, mask = s1)              :870 */
    LGPU_OP_COUNT
};

/* ---- context ------------------------------------------------------------------------------------ */
/* ring.NewRing x2 + ring.NewBasisExtender + ring.NewDecomposer (ring/ring.go:258-322,
 * ring/basis_extension.go:52-87,318-377): builds every per-prime constant (BRed/MRed constants, NInv,
 * RootsForward/RootsBackward with the reference's primitive-root choice ring/subring.go:161-194,
 * RescaleConstants ring/ring.go:329-346, ModUp/ModDown/Decomposer constants) and uploads them.
 * `p` may be NULL / np = 0 (ring-only use). N = 2^logN, 16 <= N <= 2^17. */
int lgpu_create(lgpu_ctx** out, int device, int logN, int ring_type,
                const uint64_t* q, int nq, const uint64_t* p, int np);
void lgpu_destroy(lgpu_ctx* ctx);
const char* lgpu_last_error(void);
const char* lgpu_version(void);
int lgpu_sync(lgpu_ctx* ctx, void* stream);

/* Table read-back (host buffers) so the Go shim / tests can check the constants against the
 * reference's exported fields (SubRing.Modulus/MRedConstant/BRedConstant/NInv/RootsForward/
 * RootsBackward, ring/subring.go:15-35, ring/ntt.go:38-44). kind: 0 = {q, qinv, bred_hi, bred_lo, ninv,
 * primitive_root} (6 words), 1 = RootsForward (NthRoot/2 words), 2 = RootsBackward,
 * 3 = RescaleConstants row `limb-1` (limb words; Q or P ring). */
int lgpu_ring_get_table(lgpu_ctx* ctx, int ring, int limb, int kind, uint64_t* host_out, size_t nwords);
/* Replaces the generated root tables of one limb by caller-supplied ones (the Go structs' own tables),
 * so constants are identical by construction. */
int lgpu_ring_set_roots(lgpu_ctx* ctx, int ring, int limb, const uint64_t* roots_fwd, const uint64_t* roots_bwd,
                        uint64_t ninv);

/* ---- device memory (mirrors BufferPool.Get / Recycle, ring/pool.go:40-61) ---------------------------- */
int lgpu_malloc(lgpu_ctx* ctx, void** dptr, size_t bytes);
int lgpu_free(lgpu_ctx* ctx, void* dptr);
int lgpu_memcpy_h2d(lgpu_ctx* ctx, void* dst, const void* src, size_t bytes, void* stream);
int lgpu_memcpy_d2h(lgpu_ctx* ctx, void* dst, const void* src, size_t bytes, void* stream);

/* ---- NTT ----------------------------------------------------------------------------------------------
 * Ring.NTT / NTTLazy / INTT / INTTLazy (ring/ntt.go:127-152). In-place (in == out) allowed.
 * lazy != 0 returns NTTLazy's representative in [0, 6q) (exact reference schedule); INTTLazy is fully
 * reduced for N >= 16 like the reference (ring/ntt.go:197-206). `batch` polynomials, `batch_stride`
 * words apart (use batch = 1, batch_stride = 0 for a single polynomial). */
int lgpu_ntt(lgpu_ctx* ctx, int ring, int level, const uint64_t* in, uint64_t* out, int lazy,
             int batch, size_t batch_stride, void* stream);
int lgpu_intt(lgpu_ctx* ctx, int ring, int level, const uint64_t* in, uint64_t* out, int lazy,
              int batch, size_t batch_stride, void* stream);
/* SubRing.NTT / NTTLazy / INTT / INTTLazy on one row (ring/subring_ops.go:235-253) */
int lgpu_subring_ntt(lgpu_ctx* ctx, int ring, int limb, const uint64_t* in, uint64_t* out, int lazy, void* stream);
int lgpu_subring_intt(lgpu_ctx* ctx, int ring, int limb, const uint64_t* in, uint64_t* out, int lazy, void* stream);

/* ---- coefficient-wise ops -------------------------------------------------------------------------------
 * Ring-level (ring/operations.go:11-392): loops over limbs 0..level. `scalars0/1` are HOST arrays with
 * one value per limb (NULL when the op takes none), e.g. AddScalar passes scalar mod q_i, MulScalar passes
 * MForm(scalar mod q_i). p1/p2 may be NULL for ops that do not read them. */
int lgpu_vecop(lgpu_ctx* ctx, int ring, int level, int opcode, const uint64_t* p1, const uint64_t* p2,
               uint64_t* p3, const uint64_t* scalars0, const uint64_t* scalars1,
               int batch, size_t batch_stride, void* stream);
/* SubRing-level (ring/subring_ops.go:6-273): one row of n words (n <= N) with immediate scalars. */
int lgpu_subring_vecop(lgpu_ctx* ctx, int ring, int limb, int opcode, const uint64_t* p1, const uint64_t* p2,
                       uint64_t* p3, uint64_t s0, uint64_t s1, int n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LATTIGO_B200_H */
