"""CPU checks that pin the oracle's restatement of the remaining Ring methods through algebraic identities (the
reference has no golden values for them): MultByMonomial == multiplication by X^k through the NTT, the
conjugate-invariant coefficient-domain automorphism == INTT o AutomorphismNTT o NTT (ring/ring_test.go style),
AutomorphismHoistedLazy followed by ModDown == AutomorphismHoisted, and the host-only C-ABI context building the
same NthRoot = 4N tables as the oracle."""
import numpy as np

from oracle import oracle as O
from tests import helpers as H

U64 = np.uint64


def test_cabi_host_tables_conjugate_invariant():
    import lattigo_b200 as lb
    for logN in (4, 8, 12):
        N = 1 << logN
        Q = H.Qi60[:2]
        ctx = lb.Context(logN, Q, device=-1, ring_type=1)
        r = O.Ring(N, Q, "ConjugateInvariant")
        for i in range(2):
            assert np.array_equal(ctx.table(0, i, 1), r.SubRings[i].RootsForward)
            assert np.array_equal(ctx.table(0, i, 2), r.SubRings[i].RootsBackward)
            assert ctx.table(0, i, 0)[4] == r.SubRings[i].NInv
        ctx.close()


def test_mult_by_monomial_is_ntt_multiplication():
    N = 64
    Q = H.Qi60[:2]
    r = O.Ring(N, Q)
    rng = np.random.default_rng(1)
    x = H.rand_poly(Q, N, rng)
    for k in (0, 1, 5, N, N + 3, 2 * N - 1, -1, -N - 2, 2 * N + 7):
        y = np.zeros_like(x)
        r.MultByMonomial(x, k, y)
        kk = k % (2 * N)
        mono = np.zeros_like(x)
        for i, q in enumerate(Q):
            mono[i, kk % N] = 1 if kk < N else q - 1
        a = np.zeros_like(x); b = np.zeros_like(x); c = np.zeros_like(x)
        r.NTT(x, a); r.NTT(mono, b); r.MForm(b, b); r.MulCoeffsMontgomery(a, b, c); r.INTT(c, c)
        yy = np.zeros_like(y)
        r.Reduce(y, yy)                      # the reference leaves q for a wrapped zero coefficient
        assert np.array_equal(yy, c), k


def test_shift_known_answer_of_the_reference():
    """ring/ring_test.go:906-919 (testShift): NewRing(16, {97}), coefficients 0..15 shifted by 3."""
    r = O.Ring(16, [97])
    x = np.arange(16, dtype=U64)[None, :].copy()
    y = np.zeros_like(x)
    r.Shift(x, 3, y)
    assert y[0].tolist() == [3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 0, 1, 2]


def test_shift_and_extend_basis():
    N = 32
    Q = H.Qi60[:2]; P = H.Pi60[:2]
    r = O.Ring(N, Q); rp = O.Ring(N, P)
    rng = np.random.default_rng(2)
    x = H.rand_poly(Q, N, rng)
    y = np.zeros_like(x)
    r.Shift(x, 5, y)
    assert np.array_equal(y[:, 0], x[:, 5]) and np.array_equal(y[:, N - 1], x[:, 4])
    r.Shift(x, -1, y)
    assert np.array_equal(y[:, 1], x[:, 0])
    sm = rng.integers(-9, 10, size=N)
    pin = np.stack([(sm % q).astype(U64) for q in Q])
    oq = np.zeros_like(pin); op = np.zeros((2, N), dtype=U64)
    O.ExtendBasisSmallNormAndCenter(r, rp, pin, 1, oq, op)
    assert np.array_equal(oq, pin)
    for j, p in enumerate(P):
        assert np.array_equal(op[j], (sm % p).astype(U64))


def test_conjugate_invariant_automorphism_consistency():
    N = 128
    Q = H.Qi60[:2]
    r = O.Ring(N, Q, "ConjugateInvariant")
    rng = np.random.default_rng(3)
    x = H.rand_poly(Q, N, rng)
    for gal in (5, 25, 125, pow(5, 33, 4 * N)):
        a = np.zeros_like(x)
        r.NTT(x, a)
        idx = r.AutomorphismNTTIndex(gal)
        b = np.zeros_like(x)
        r.AutomorphismNTTWithIndex(a, idx, b)
        r.INTT(b, b)
        c = np.zeros_like(x)
        r.Automorphism(x, gal, c)
        cc = np.zeros_like(c)
        r.Reduce(c, cc)
        assert np.array_equal(b, cc), gal


def test_automorphism_hoisted_lazy_then_moddown():
    logN = 6
    q, p = O.gen_moduli(logN + 1, [56, 45, 45, 45], [55, 55])
    params = O.Parameters(logN, q, p)
    N = params.N()
    rng = np.random.default_rng(4)
    ev = O.Evaluator(params)
    levelQ, levelP = params.MaxLevelQ(), params.MaxLevelP()
    evk = H.random_gadget_ciphertext(params, levelQ, levelP, rng)
    n = params.BaseRNSDecompositionVectorSize(levelQ, levelP)
    ct = [H.rand_poly(q, N, rng) for _ in range(2)]
    dq = [np.zeros((levelQ + 1, N), dtype=U64) for _ in range(n)]; dp = [np.zeros((levelP + 1, N), dtype=U64) for _ in range(n)]
    ev.DecomposeNTT(levelQ, levelP, levelP + 1, ct[1].copy(), True, dq, dp)
    gal = params.GaloisElement(2)
    want = [np.zeros((levelQ + 1, N), dtype=U64) for _ in range(2)]
    ev.AutomorphismHoisted(levelQ, ct, dq, dp, gal, evk, want)
    lq = [np.zeros((levelQ + 1, N), dtype=U64) for _ in range(2)]; lp = [np.zeros((levelP + 1, N), dtype=U64) for _ in range(2)]
    ev.AutomorphismHoistedLazy(levelQ, ct, dq, dp, gal, evk, lq, lp)
    got = [np.zeros((levelQ + 1, N), dtype=U64) for _ in range(2)]
    ev.ModDown(levelQ, levelP, lq, lp, got)
    # ct0 * P + gadget, divided by P: the rounding of the P-scaled ct0 term is exact, so both routes agree
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
