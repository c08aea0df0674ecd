"""Anchoring of oracle/blindrot.py (core/rgsw/blindrot/evaluator.go:144-283). The reference tests blind rotation end to end by decrypting
(blindrot_test.go:60-160); here the restated control flow is pinned the same way on a small ring: with blind-rotation keys RGSW(X^{s_j}) and the
window's automorphism keys, BlindRotateCore turns an accumulator encrypting h(X) into one encrypting h(X^t) * X^{<a, s>} with t = -g^{-1} mod 2N
(the composition of all the automorphisms of Algorithm 3; Evaluate feeds it f(X^{-g}) X^{-g b}, which gives f(X) X^{b + <a, s>}: evaluator.go:107-124)."""
import numpy as np

from oracle import blindrot as BR
from oracle import oracle as O
from tests import helpers as H

U64 = np.uint64
LOGN = 5


def _auto_int(poly, g, N):
    out = [0] * N
    for i, c in enumerate(poly):
        e = (i * g) % (2 * N)
        if e < N: out[e] += int(c)
        else: out[e - N] -= int(c)
    return out


def _mul_monomial(poly, e, N):
    out = [0] * N
    for i, c in enumerate(poly):
        k = (i + e) % (2 * N)
        if k < N: out[k] += int(c)
        else: out[k - N] -= int(c)
    return out


def setup(seed, n_lwe=6, window=3, lq=(50, 45), lp=(50,), pw2=0):
    q, p = O.gen_moduli(LOGN + 1, list(lq), list(lp))
    params = O.Parameters(LOGN, q, p)
    N = params.N()
    rng = np.random.default_rng(seed)
    s = rng.integers(-1, 2, N)
    s_lwe = rng.integers(-1, 2, n_lwe)
    brk = []
    for j in range(n_lwe):
        mono = np.zeros(N, dtype=np.int64)
        if s_lwe[j] == 0: mono[0] = 1
        elif s_lwe[j] == 1: mono[1] = 1
        else: mono[N - 1] = -1                                   # X^{-1} = -X^{N-1}
        brk.append(H.gen_rgsw(params, list(mono), list(s), rng, pw2=pw2))
    ring1 = O.Ring(N, [q[0]])
    sp = H.small_poly_rns(s, [q[0]])
    gks = {}
    for g in [params.GaloisElement(k) for k in range(1, window + 1)] + [2 * N - 5]:
        ginv = pow(g, -1, 2 * N)
        row = np.empty_like(sp); ring1.Automorphism(sp, ginv, row)
        s_out = [int(v) if int(v) < q[0] // 2 else int(v) - q[0] for v in row[0]]
        gks[g] = H.gen_switching_key(params, s, s_out, rng, pw2=pw2)
    a = rng.integers(0, N, n_lwe) * 2 + 1                        # odd residues modulo 2N
    a[0] = 0                                                     # zero is allowed too (Go's map lookup yields k = 0)
    return params, rng, s, s_lwe, brk, gks, a, q


def test_blind_rotate_core_decrypts_to_the_rotated_accumulator():
    window = 3
    params, rng, s, s_lwe, brk, gks, a, q = setup(1, window=window)
    N = params.N()
    levelQ = len(q) - 1
    ringQ = params.ringQ
    h = rng.integers(-3, 4, N) * (1 << 30)
    acc = [H.small_poly_rns(h, q), np.zeros((levelQ + 1, N), dtype=U64)]
    ringQ.NTT(acc[0], acc[0])
    BR.Evaluator(params, window).BlindRotateCore(a, acc, brk, gks)
    sn = np.empty((levelQ + 1, N), dtype=U64); ringQ.NTT(H.small_poly_rns(s, q), sn)
    dec = np.stack([np.array([(int(x) + int(y) * int(t)) % int(m) for x, y, t in zip(acc[0][l], acc[1][l], sn[l])], dtype=U64) for l, m in enumerate(q)])
    ringQ.INTT(dec, dec)
    Q = ringQ.ModulusAtLevel[levelQ]
    got = [v if v < Q // 2 else v - Q for v in ringQ.PolyToBigint(dec)]
    t = (-pow(5, -1, 2 * N)) % (2 * N)
    # a zero entry of the mask has no discrete log; Go's map lookup returns 0, i.e. the entry is handled as g^0 = 1 (evaluator.go:264-276)
    e = int(sum((int(x) if x else 1) * int(y) for x, y in zip(a, s_lwe))) % (2 * N)
    want = _mul_monomial(_auto_int(h, t, N), e, N)
    err = max(abs(g - w) for g, w in zip(got, want))
    assert np.log2(err + 1) < 14, np.log2(err + 1)               # message at 2^30; observed noise ~2^6 (N/2 key switches and 6 external products)


def test_discrete_log_sets_and_missing_key():
    N = 32
    dl = BR.galois_element_inverse_map(N)
    assert dl[1] == 0 and dl[5] == 1 and dl[2 * N - 5] == -1 and len(dl) == N
    sets = BR.discrete_log_sets([5, 0, 1, 2 * N - 25], N)
    assert sets == {1: [0], 0: [1, 2], -2: [3]}
    params, rng, s, s_lwe, brk, gks, a, q = setup(2, window=2)
    del gks[2 * N * 0 + params.GaloisElement(1)]
    acc = [H.rand_poly(q, params.N(), rng), H.rand_poly(q, params.N(), rng)]
    try:
        BR.Evaluator(params, 2).BlindRotateCore(a, acc, brk, gks)
    except KeyError as ex:
        assert "GaloisKey" in str(ex)
    else:
        raise AssertionError("expected a missing-key error")
