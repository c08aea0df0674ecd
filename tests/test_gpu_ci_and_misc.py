"""GPU parity (bit-exact) against the oracle for: the conjugate-invariant transforms (ring/ntt.go:716-1311) and the
CI automorphisms; the coefficient-moving Ring methods (Shift, MultByMonomial, MapSmallDimensionToLargerDimensionNTT,
ExtendBasisSmallNormAndCenter); the *RNSScalar / EvalPolyScalar / MulByVector wrappers; AutomorphismHoistedLazy.
Runs only on the B200 box (-m gpu)."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu
U64 = np.uint64


def _lb():
    import lattigo_b200 as lb
    return lb


@pytest.mark.parametrize("logN", [4, 5, 7, 10, 12, 13, 15, 16])
def test_conjugate_invariant_ntt(logN):
    """NTTConjugateInvariant[Lazy] / INTTConjugateInvariant[Lazy]: exact representatives, batches, in place, round trip.
    (Qi60 primes are 1 mod 2^18, so they are valid for NthRoot = 4N up to N = 2^16.)"""
    lb = _lb()
    N = 1 << logN
    Q = H.Qi60[:3]
    ctx = lb.Context(logN, Q, ring_type=1)
    ring = O.Ring(N, Q, "ConjugateInvariant")
    # constants derived on the device side agree with the oracle's (NthRoot = 4N tables)
    for i in range(len(Q)):
        assert np.array_equal(ctx.table(0, i, 1), ring.SubRings[i].RootsForward)
        assert np.array_equal(ctx.table(0, i, 2), ring.SubRings[i].RootsBackward)
    rng = np.random.default_rng(500 + logN)
    batch = 2
    x = np.stack([H.rand_poly(Q, N, rng) for _ in range(batch)])
    want = np.empty_like(x); want_lazy = np.empty_like(x); want_inv = np.empty_like(x); want_inv_lazy = np.empty_like(x)
    for b in range(batch):
        ring.NTT(x[b], want[b]); ring.NTTLazy(x[b], want_lazy[b]); ring.INTT(x[b], want_inv[b]); ring.INTTLazy(x[b], want_inv_lazy[b])
    d = ctx.to_device(x)
    out = ctx.ringQ.NewPoly(batch)
    ctx.ringQ.NTT(d, out)
    assert np.array_equal(ctx.to_host(out), want)
    ctx.ringQ.NTTLazy(d, out)
    assert np.array_equal(ctx.to_host(out), want_lazy)
    ctx.ringQ.INTT(d, out)
    assert np.array_equal(ctx.to_host(out), want_inv)
    ctx.ringQ.INTTLazy(d, out)
    assert np.array_equal(ctx.to_host(out), want_inv_lazy)        # [0, 2q) representative (MRedLazy by NInv)
    e = d.clone()
    ctx.ringQ.NTT(e, e)
    assert np.array_equal(ctx.to_host(e), want)
    ctx.ringQ.INTT(e, e)
    assert np.array_equal(ctx.to_host(e), x)
    # single row through the SubRing entry point
    sr = ctx.ringQ.SubRings[1]
    o1 = ctx.new_poly(1)
    sr.NTT(d[0, 1:2], o1)
    assert np.array_equal(ctx.to_host(o1)[0], want[0, 1])
    ctx.close()


def test_conjugate_invariant_automorphisms():
    lb = _lb()
    logN = 9
    N = 1 << logN
    Q = H.Qi60[:2]
    ctx = lb.Context(logN, Q, ring_type=1)
    ring = O.Ring(N, Q, "ConjugateInvariant")
    rng = np.random.default_rng(77)
    x = H.rand_poly(Q, N, rng)
    d = ctx.to_device(x)
    for gal in (5, 25, pow(5, 77, 4 * N), pow(5, 4 * N - 3, 4 * N)):   # X -> X^-1 (gal = -1) is the identity on this ring and has no NTT index
        want = np.empty_like(x)
        ring.Automorphism(x, gal, want)
        out = ctx.ringQ.NewPoly()
        ctx.ringQ.Automorphism(d, gal, out)
        assert np.array_equal(ctx.to_host(out), want), gal
        idx = ring.AutomorphismNTTIndex(gal)
        didx = ctx.ringQ.AutomorphismNTTIndex(gal)
        assert np.array_equal(ctx.to_host(didx), idx), gal
        want_n = np.empty_like(x)
        ring.AutomorphismNTTWithIndex(x, idx, want_n)
        ctx.ringQ.AutomorphismNTT(d, gal, out)
        assert np.array_equal(ctx.to_host(out), want_n), gal
    ctx.close()


def test_shift_monomial_map_extend():
    lb = _lb()
    logN = 10
    N = 1 << logN
    Q = H.Qi60[:3]; P = H.Pi60[:2]
    ctx = lb.Context(logN, Q, P)
    ring = O.Ring(N, Q); ringP = O.Ring(N, P)
    rng = np.random.default_rng(8)
    batch = 2
    x = np.stack([H.rand_poly(Q, N, rng) for _ in range(batch)])
    x[0, 0, :7] = 0                                            # zero coefficients exercise the literal `q - 0 = q`
    x[1, 2, N - 5:] = 0
    d = ctx.to_device(x)
    out = ctx.ringQ.NewPoly(batch)
    for k in (0, 1, 17, N - 1, N, N + 3, 2 * N - 1, -1, -N, -N - 9, -2 * N + 1, 2 * N, 3 * N + 5):
        want = np.zeros_like(x)
        for b in range(batch):
            ring.MultByMonomial(x[b], k, want[b])
        ctx.ringQ.MultByMonomial(d, k, out)
        assert np.array_equal(ctx.to_host(out), want), k
        e = d.clone()
        ctx.ringQ.MultByMonomial(e, k, e)                      # in place (the reference goes through a temporary)
        assert np.array_equal(ctx.to_host(e), want), k
    for k in (0, 1, 5, N - 1, N, N + 2, -1, -3 * N - 4):
        want = np.zeros_like(x)
        for b in range(batch):
            ring.Shift(x[b], k, want[b])
        ctx.ringQ.Shift(d, k, out)
        assert np.array_equal(ctx.to_host(out), want), k
        e = d.clone()
        ctx.ringQ.Shift(e, k, e)
        assert np.array_equal(ctx.to_host(e), want), k
    # level < max touches only rows 0..level
    lvl = ctx.ringQ.AtLevel(1)
    o2 = ctx.to_device(np.full((3, N), 7, dtype=U64))
    lvl.MultByMonomial(d[0], 3, o2)
    w = np.zeros((3, N), dtype=U64)
    ring.AtLevel(1).MultByMonomial(x[0], 3, w)
    got = ctx.to_host(o2)
    assert np.array_equal(got[:2], w[:2]) and np.all(got[2] == 7)
    # MapSmallDimensionToLargerDimensionNTT
    small = H.rand_poly(Q, 64, rng)
    large = np.zeros((3, N), dtype=U64)
    O.MapSmallDimensionToLargerDimensionNTT(small, large)
    dl = ctx.ringQ.NewPoly()
    lb.ring.MapSmallDimensionToLargerDimensionNTT(ctx, ctx.to_device(small), dl)
    assert np.array_equal(ctx.to_host(dl), large)
    # ExtendBasisSmallNormAndCenter: small-norm input (ternary / small gaussian-like) on row 0
    sm = rng.integers(-40, 41, size=(batch, N))
    pin = np.stack([np.stack([(sm[b] % q).astype(U64) for q in Q]) for b in range(batch)])
    from lattigo_b200.ringqp import RingQP
    rqp = RingQP(ctx)
    dq = ctx.to_device(pin); oq = ctx.ringQ.NewPoly(batch); op = ctx.ringP.NewPoly(batch)
    rqp.ExtendBasisSmallNormAndCenter(dq, len(P) - 1, oq, op)
    for b in range(batch):
        wq = np.zeros((3, N), dtype=U64); wp = np.zeros((2, N), dtype=U64)
        O.ExtendBasisSmallNormAndCenter(ring, ringP, pin[b], len(P) - 1, wq, wp)
        assert np.array_equal(ctx.to_host(oq)[b], wq) and np.array_equal(ctx.to_host(op)[b], wp)
        assert np.array_equal(wp[0], (sm[b] % P[0]).astype(U64))        # sanity: it is the centred lift
    ctx.close()


def test_rns_scalar_wrappers():
    lb = _lb()
    logN = 8
    N = 1 << logN
    Q = H.Qi60[:3]
    ctx = lb.Context(logN, Q)
    ring = O.Ring(N, Q)
    rng = np.random.default_rng(9)
    x = H.rand_poly(Q, N, rng); y = H.rand_poly(Q, N, rng)
    s0 = [int(rng.integers(0, q)) for q in Q]; s1 = [int(rng.integers(0, q)) for q in Q]
    d = ctx.to_device(x)
    for name in ("AddDoubleRNSScalar", "SubDoubleRNSScalar", "MulDoubleRNSScalar", "MulDoubleRNSScalarThenAdd"):
        want = y.copy(); out = ctx.to_device(y)
        getattr(ring, name)(x, s0, s1, want)
        getattr(ctx.ringQ, name)(d, s0, s1, out)
        assert np.array_equal(ctx.to_host(out), want), name
    want = y.copy(); out = ctx.to_device(y)
    ring.MulRNSScalarMontgomery(x, s0, want); ctx.ringQ.MulRNSScalarMontgomery(d, s0, out)
    assert np.array_equal(ctx.to_host(out), want)
    big = (1 << 100) + 12345
    want = y.copy(); out = ctx.to_device(y)
    ring.MulScalarThenSub(x, big, want); ctx.ringQ.MulScalarThenSub(d, big, out)
    assert np.array_equal(ctx.to_host(out), want)
    # EvalPolyScalar: Horner over three coefficient polynomials
    pols = [H.rand_poly(Q, N, rng) for _ in range(3)]
    want = np.zeros_like(x); out = ctx.ringQ.NewPoly()
    ring.EvalPolyScalar(pols, 0xDEADBEEF, want)
    ctx.ringQ.EvalPolyScalar([ctx.to_device(p) for p in pols], 0xDEADBEEF, out)
    assert np.array_equal(ctx.to_host(out), want)
    # MulByVectorMontgomery[ThenAddLazy]: one vector against every limb
    vec = rng.integers(0, min(Q), size=N, dtype=np.int64).astype(U64)
    dv = ctx.to_device(vec)
    want = np.zeros_like(x); out = ctx.ringQ.NewPoly()
    ring.MulByVectorMontgomery(x, vec, want); ctx.ringQ.MulByVectorMontgomery(d, dv, out)
    assert np.array_equal(ctx.to_host(out), want)
    want = y.copy(); out = ctx.to_device(y)
    ring.MulByVectorMontgomeryThenAddLazy(x, vec, want); ctx.ringQ.MulByVectorMontgomeryThenAddLazy(d, dv, out)
    assert np.array_equal(ctx.to_host(out), want)
    ctx.close()


def test_automorphism_hoisted_lazy():
    lb = _lb()
    logN = 8
    q, p = O.gen_moduli(logN + 1, [56, 45, 45, 45, 45, 45, 45], [55, 55, 55])
    ctx = lb.Context(logN, q, p)
    params = O.Parameters(logN, q, p)
    N = params.N()
    rng = np.random.default_rng(61)
    ev_o = O.Evaluator(params); ev = lb.Evaluator(ctx)
    levelP = params.MaxLevelP()
    evk_o = H.random_gadget_ciphertext(params, params.MaxLevelQ(), levelP, rng)
    evk = lb.GadgetCiphertext(ctx, evk_o.data, evk_o.LevelQ(), evk_o.LevelP())
    gal = params.GaloisElement(3)
    for levelQ in (params.MaxLevelQ(), 2):
        n = params.BaseRNSDecompositionVectorSize(levelQ, levelP)
        ct = np.stack([H.rand_poly(q[: levelQ + 1], N, rng) for _ in range(2)])
        dq = [np.zeros((levelQ + 1, N), dtype=U64) for _ in range(n)]; dp = [np.zeros((levelP + 1, N), dtype=U64) for _ in range(n)]
        ev_o.DecomposeNTT(levelQ, levelP, levelP + 1, ct[1].copy(), True, dq, dp)
        wQ = [np.zeros((levelQ + 1, N), dtype=U64) for _ in range(2)]; wP = [np.zeros((levelP + 1, N), dtype=U64) for _ in range(2)]
        ev_o.AutomorphismHoistedLazy(levelQ, [ct[0], ct[1]], dq, dp, gal, evk_o, wQ, wP)
        dct = ctx.to_device(ct)
        dec = ev.DecomposeNTT(levelQ, levelP, levelP + 1, dct[1].contiguous(), True)
        oQ = [ctx.new_poly(levelQ + 1), ctx.new_poly(levelQ + 1)]; oP = [ctx.new_poly(levelP + 1), ctx.new_poly(levelP + 1)]
        ev.AutomorphismHoistedLazy(levelQ, dct, dec, gal, evk, oQ, oP)
        for c in range(2):
            assert np.array_equal(ctx.to_host(oQ[c]), wQ[c]), (levelQ, c)
            assert np.array_equal(ctx.to_host(oP[c]), wP[c]), (levelQ, c)
    ctx.close()
