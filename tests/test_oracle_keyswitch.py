"""Algebraic anchoring of the oracle's key-switch family, for which the reference has no bit-level
vectors (only noise bounds: core/rlwe/rlwe_test.go:666-779 testGadgetProduct, :897-1070 testAutomorphism).
d0 + d1*s_out - c*s_in must be small, with the same bound shape the reference uses
(log2(noise) <= logN + bpw2 margin). Covers the three code paths of core/rlwe/test_params.go:17-49."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import helpers as H

U64 = np.uint64
LOGN = 7   # small ring: these run in pure-Python setup + C arithmetic


def _mods_for(logN):
    # NTT-friendly primes for a small ring with the same bit sizes as core/rlwe/test_params.go
    q, p = O.gen_moduli(logN + 1, [45, 35, 35, 35, 35], [50, 50])
    return q, p


@pytest.mark.parametrize("case", ["multipleP", "singleP_pw2", "noP_pw2", "singleP_nopw2"])
def test_gadget_product_decrypts(case):
    q, p = _mods_for(LOGN)
    pw2 = 0
    if case == "singleP_pw2":
        p, pw2 = p[:1], 16
    elif case == "noP_pw2":
        p, pw2 = [], 2
    elif case == "singleP_nopw2":
        p = p[:1]
    params = O.Parameters(LOGN, q, p)
    N = params.N()
    rng = np.random.default_rng(42)
    s_out = rng.integers(-1, 2, N)
    s_in = rng.integers(-1, 2, N)
    evk = H.gen_switching_key(params, s_in, s_out, rng, pw2=pw2)
    ev = O.Evaluator(params)
    for levelQ in (params.MaxLevelQ(), 1, 0):
        cx = H.rand_poly(q[: levelQ + 1], N, rng)
        ct = [np.zeros((levelQ + 1, N), dtype=U64) for _ in range(2)]
        ev.GadgetProduct(levelQ, cx, evk, ct)
        noise = H.keyswitch_noise_log2(params, levelQ, cx, ct, s_in, s_out)
        bound = LOGN + 12 + (pw2 if pw2 else 0) + (35 if case == "singleP_nopw2" and False else 0)
        if case == "noP_pw2":
            bound = LOGN + pw2 + 12
        assert noise < bound, (case, levelQ, noise, bound)
        for l, m in enumerate(q[: levelQ + 1]):
            assert int(ct[0][l].max()) < m and int(ct[1][l].max()) < m   # canonical outputs


def test_hoisted_equals_plain_and_automorphism_decrypts():
    q, p = _mods_for(LOGN)
    params = O.Parameters(LOGN, q, p)
    N = params.N()
    rng = np.random.default_rng(1)
    ev = O.Evaluator(params)
    s = rng.integers(-1, 2, N)
    galEl = params.GaloisElement(3)
    # Galois key (core/rlwe/keygenerator.go GenGaloisKey): encrypts s under pi_{g^-1}(s); Automorphism
    # key-switches ct[1] and then permutes (core/rlwe/evaluator_automorphism.go:41-47), so the output
    # decrypts under s to pi_g(m).
    ginv = pow(galEl, -1, 2 * N)
    ring1 = O.Ring(N, [q[0]])
    sp = H.small_poly_rns(s, [q[0]])
    s_out_row = np.empty_like(sp); ring1.Automorphism(sp, ginv, s_out_row)
    s_out = [int(v) if int(v) < q[0] // 2 else int(v) - q[0] for v in s_out_row[0]]
    gk = H.gen_switching_key(params, s, s_out, rng)
    levelQ = params.MaxLevelQ()
    levelP = params.MaxLevelP()
    c = [H.rand_poly(q, N, rng) for _ in range(2)]
    # plain vs hoisted gadget product: identical bits
    ct_a = [np.zeros((levelQ + 1, N), dtype=U64) for _ in range(2)]
    ev.GadgetProduct(levelQ, c[1], gk, ct_a)
    n = params.BaseRNSDecompositionVectorSize(levelQ, levelP)
    dq = [np.zeros((levelQ + 1, N), dtype=U64) for _ in range(n)]
    dp = [np.zeros((levelP + 1, N), dtype=U64) for _ in range(n)]
    ev.DecomposeNTT(levelQ, levelP, levelP + 1, c[1], True, dq, dp)
    ct_b = [np.zeros((levelQ + 1, N), dtype=U64) for _ in range(2)]
    ev.GadgetProductHoisted(levelQ, dq, dp, gk, ct_b)
    assert np.array_equal(ct_a[0], ct_b[0]) and np.array_equal(ct_a[1], ct_b[1])
    # Automorphism: out decrypts (under s) to pi_g(c0 + c1*s)
    out = [np.zeros((levelQ + 1, N), dtype=U64) for _ in range(2)]
    ev.Automorphism(c, galEl, gk, out)
    ringQ = params.ringQ
    sn = np.empty((levelQ + 1, N), dtype=U64); ringQ.NTT(H.small_poly_rns(s, q), sn)
    def dec(ct):
        acc = np.stack([np.array([(int(a) + int(b) * int(t)) % int(m) for a, b, t in zip(ct[0][l], ct[1][l], sn[l])], dtype=U64) for l, m in enumerate(q)])
        back = np.empty_like(acc); ringQ.INTT(acc, back)
        return back
    m_in = dec(c)
    want = np.empty_like(m_in); ringQ.Automorphism(m_in, galEl, want)
    got = dec(out)
    Q = ringQ.ModulusAtLevel[levelQ]
    diff = [(a - b) % Q for a, b in zip(ringQ.PolyToBigint(got), ringQ.PolyToBigint(want))]
    m = max(min(v, Q - v) for v in diff)
    assert np.log2(float(m) + 1) < LOGN + 12
