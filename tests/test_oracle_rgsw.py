"""Anchoring of oracle/rgsw.py (core/rgsw/evaluator.go:39-283). The reference has no bit-level vectors for
the external product (core/rgsw/rgsw_test.go:60-110 checks the noise of the decrypted product), so the restatement is pinned
the same way: a real RLWE encryption of m0 times a real RGSW encryption of m1 must decrypt to m0 * m1 up to noise, on each of
the three code paths; and the multiple-P path must equal the sum of the two (already pinned) lazy gadget products."""
import numpy as np

from oracle import oracle as O
from oracle import rgsw as RG
from tests import helpers as H

U64 = np.uint64


def _negacyclic(a, b, N):
    out = [0] * N
    for i, x in enumerate(a):
        if x == 0: continue
        for j, y in enumerate(b):
            e = i + j
            if e < N: out[e] += int(x) * int(y)
            else: out[e - N] -= int(x) * int(y)
    return out


def _case(logN, q, p, pw2, seed, log_delta):
    params = O.Parameters(logN, q, p)
    N = params.N()
    rng = np.random.default_rng(seed)
    s = rng.integers(-1, 2, N)
    m1 = np.zeros(N, dtype=np.int64); m1[rng.integers(0, N, 3)] = rng.integers(-1, 2, 3); m1[0] = 1     # small RGSW message
    m0 = rng.integers(-8, 9, N)
    rg = H.gen_rgsw(params, list(m1), list(s), rng, pw2=pw2)
    ringQ = params.ringQ
    levelQ = params.MaxLevelQ()
    a = H.rand_poly(q, N, rng)
    sn = np.empty((levelQ + 1, N), dtype=U64); ringQ.NTT(H.small_poly_rns(s, q), sn)
    me = H.small_poly_rns(m0 * (1 << log_delta) + np.rint(rng.normal(0, 3.2, N)).astype(np.int64), q)
    ringQ.NTT(me, me)
    c0 = np.stack([np.array([(int(x) - int(y) * int(t)) % int(mod) for x, y, t in zip(me[l], a[l], sn[l])], dtype=U64) for l, mod in enumerate(q)])
    ct = [c0, a]
    out = [np.zeros((levelQ + 1, N), dtype=U64) for _ in range(2)]
    RG.Evaluator(params).ExternalProduct(ct, rg, out)
    acc = np.stack([np.array([(int(x) + int(y) * int(t)) % int(mod) for x, y, t in zip(out[0][l], out[1][l], sn[l])], dtype=U64) for l, mod in enumerate(q)])
    back = np.empty_like(acc); ringQ.INTT(acc, back)
    Q = ringQ.ModulusAtLevel[levelQ]
    got = [v if v < Q // 2 else v - Q for v in ringQ.PolyToBigint(back)]
    want = [v << log_delta for v in _negacyclic(m0, m1, N)]
    err = max(abs(g - w) for g, w in zip(got, want))
    for l, mod in enumerate(q):
        assert int(out[0][l].max()) < mod and int(out[1][l].max()) < mod
    return float(np.log2(err + 1)), params, ct, rg, out


def test_external_product_multiple_p_decrypts_and_equals_two_gadget_products():
    q, p = O.gen_moduli(7, [45, 35, 35, 35], [50, 50])
    err, params, ct, rg, out = _case(6, q, p, 0, 1, 20)
    assert err < 16, err
    ev = O.Evaluator(params)
    N, levelQ, levelP = params.N(), 3, 1
    accQ = [np.zeros((2, levelQ + 1, N), dtype=U64) for _ in range(2)]; accP = [np.zeros((2, levelP + 1, N), dtype=U64) for _ in range(2)]
    for k in range(2):
        ev.GadgetProductLazy(levelQ, ct[k].copy(), rg[k], [accQ[k][0], accQ[k][1]], [accP[k][0], accP[k][1]])
    sQ = np.zeros((2, levelQ + 1, N), dtype=U64); sP = np.zeros((2, levelP + 1, N), dtype=U64)
    for c in range(2):
        params.ringQ.Add(accQ[0][c], accQ[1][c], sQ[c]); params.ringP.Add(accP[0][c], accP[1][c], sP[c])
    want = [np.zeros((levelQ + 1, N), dtype=U64) for _ in range(2)]
    ev.ModDown(levelQ, levelP, [sQ[0], sQ[1]], [sP[0], sP[1]], want)
    assert np.array_equal(out[0], want[0]) and np.array_equal(out[1], want[1])
    # in place (the only way the reference's own callers use it: rgsw_test.go:84, blindrot/evaluator.go:212)
    inpl = [ct[0].copy(), ct[1].copy()]
    RG.Evaluator(params).ExternalProduct(inpl, rg, inpl)
    assert np.array_equal(inpl[0], out[0]) and np.array_equal(inpl[1], out[1])


def test_external_product_single_p_bit_decomposition_decrypts():
    q, p = O.gen_moduli(7, [45, 35, 35], [50])
    err, *_ = _case(6, q, p, 12, 2, 20)
    assert err < 22, err                                   # message at 2^20+, digits of 12 bits: noise ~ 2^12 * N * sigma


def test_external_product_no_p_paths_decrypt():
    q, _ = O.gen_moduli(7, [50, 40], [])
    err, *_ = _case(6, q, [], 10, 3, 24)
    assert err < 20, err
    q32, _ = O.gen_moduli(7, [27], [])                      # externalProduct32Bit: one modulus below 2^29, lazy 64-bit sums
    assert q32[0] >> 29 == 0
    err, *_ = _case(6, q32, [], 7, 4, 16)
    assert err < 14, err
