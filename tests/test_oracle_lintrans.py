"""Algebraic anchoring of oracle/lintrans.py (hoisted rotations + BSGS linear transformation).
The reference has no bit-level vectors for circuits/common/lintrans; as for the key-switch family, the restatement is
pinned by decryption: with real Galois keys, the output ciphertext must decrypt (under s) to
sum_k diag_k * pi_{5^k}(m) up to key-switch noise, for the naive (single hoisting) and the BSGS (double hoisting)
evaluators (for BSGS with the diagonals' giant-step pre-rotation taken into account)."""
import numpy as np

from oracle import lintrans as LT
from oracle import oracle as O
from tests import helpers as H

U64 = np.uint64
LOGN = 6


def _setup(seed, rots, N1):
    q, p = O.gen_moduli(LOGN + 1, [45, 35, 35, 35], [50, 50])
    params = O.Parameters(LOGN, q, p)
    N = params.N()
    rng = np.random.default_rng(seed)
    s = rng.integers(-1, 2, N)
    ringQ = params.ringQ; ringP = params.ringP
    levelQ, levelP = params.MaxLevelQ(), params.MaxLevelP()
    ring1 = O.Ring(N, [q[0]])
    sp = H.small_poly_rns(s, [q[0]])
    # Galois keys for every rotation the evaluator can ask for (baby steps, giant steps, plain diagonals)
    slots = N >> 1
    need = set()
    if N1:
        index, rotN1, rotN2 = LT.bsgs_index(rots, slots, N1)
        need |= {r for r in rotN1 if r} | {r for r in rotN2 if r}
    else:
        need |= {r & (slots - 1) for r in rots if r & (slots - 1)}
    keys = {}
    for r in sorted(need):
        galEl = params.GaloisElement(r)
        ginv = pow(galEl, -1, 2 * N)
        row = np.empty_like(sp); ring1.Automorphism(sp, ginv, row)
        s_out = [int(v) if int(v) < q[0] // 2 else int(v) - q[0] for v in row[0]]
        keys[galEl] = H.gen_switching_key(params, s, s_out, rng)
    # diagonals: small polynomials, NTT + Montgomery form on Q and P (what LinearTransformation.Vec stores)
    vec, diag_coeffs = {}, {}
    for r in rots:
        d = rng.integers(-4, 5, N)
        diag_coeffs[r] = d
        dq = H.small_poly_rns(d, q); dp = H.small_poly_rns(d, p)
        ringQ.NTT(dq, dq); ringQ.MForm(dq, dq)
        ringP.NTT(dp, dp); ringP.MForm(dp, dp)
        vec[r] = (dq, dp)
    # a fresh encryption of a message with ~20-bit coefficients: (c0, c1) = (-a s + m + e, a), NTT domain
    m = rng.integers(-(1 << 20), 1 << 20, N)
    a = H.rand_poly(q, N, rng)
    sn = np.empty((levelQ + 1, N), dtype=U64); ringQ.NTT(H.small_poly_rns(s, q), sn)
    me = H.small_poly_rns(m + np.rint(rng.normal(0, 3.2, N)).astype(np.int64), q)
    ringQ.NTT(me, me)
    c0 = np.stack([np.array([(int(x) - int(y) * int(t)) % int(mod) for x, y, t in zip(me[l], a[l], sn[l])], dtype=U64) for l, mod in enumerate(q)])
    return params, rng, s, sn, keys, vec, diag_coeffs, m, [c0, a], levelQ, levelP


def _decrypt_centered(params, ct, sn, levelQ):
    ringQ = params.ringQ.AtLevel(levelQ)
    q = params.qi[: levelQ + 1]
    acc = np.stack([np.array([(int(a) + int(b) * int(t)) % int(mod) for a, b, t in zip(ct[0][l], ct[1][l], sn[l])], dtype=U64) for l, mod in enumerate(q)])
    back = np.empty_like(acc); ringQ.INTT(acc, back)
    Q = ringQ.ModulusAtLevel[levelQ]
    return [v if v < Q // 2 else v - Q for v in ringQ.PolyToBigint(back)]


def _auto(params, poly, k):
    """pi_{5^k} on an integer polynomial: X^i -> X^(i 5^k) in Z[X]/(X^N+1)."""
    N = params.N()
    g = params.GaloisElement(k)
    out = [0] * N
    for i in range(N):
        e = (i * g) % (2 * N)
        if e < N: out[e] += int(poly[i])
        else: out[e - N] -= int(poly[i])
    return out


def _expected(params, m, diag_coeffs, N1=0):
    """sum_k D_k * pi_{5^k}(m) in Z[X]/(X^N+1) with exact integers. Naive evaluator: D_k = diag_k. BSGS evaluator: the
    stored diagonal of index k = j + i (giant step j, baby step i) is multiplied BEFORE the giant-step rotation
    (lintrans_evaluator.go:349-431), so D_k = pi_{5^j}(diag_k) -- the reference's encoder pre-rotates the diagonals by -j
    for exactly that reason (lintrans.go:242-262: rot = -j)."""
    N = params.N()
    out = [0] * N
    for r, d0 in diag_coeffs.items():
        d = _auto(params, d0, (r // N1) * N1) if N1 else [int(v) for v in d0]
        pm = _auto(params, m, r)
        for i in range(N):
            if d[i] == 0: continue
            for j in range(N):
                e = i + j
                if e < N: out[e] += int(d[i]) * pm[j]
                else: out[e - N] -= int(d[i]) * pm[j]
    return out


def test_bsgs_index_matches_reference_example():
    index, rotN1, rotN2 = LT.bsgs_index([0, 1, 2, 3, 15, 16, 17, 31], 32, 4)
    assert index == {0: [0, 1, 2, 3], 12: [3], 16: [0, 1], 28: [3]} and rotN1 == [0, 12, 16, 28] and rotN2 == [0, 1, 2, 3]


def test_naive_and_bsgs_decrypt_to_the_linear_map():
    rots = [0, 1, 2, 5, 6, 9]
    for N1 in (0, 4):
        params, rng, s, sn, keys, vec, diag_coeffs, m, ct, levelQ, levelP = _setup(7, rots, N1)
        N = params.N()
        ev = LT.Evaluator(params, keys)
        lt = LT.LinearTransformation(vec, levelQ, levelP, LOGN - 1, N1)
        out = [np.zeros((levelQ + 1, N), dtype=U64) for _ in range(2)]
        ev.EvaluateMany(ct, [lt], [out])
        got = _decrypt_centered(params, out, sn, levelQ)
        want = _expected(params, m, diag_coeffs, N1)
        err = max(abs(a - b) for a, b in zip(got, want))
        # message terms are ~2^20 * 4 * N * len(rots); the residual must be key-switch / encryption noise only
        assert np.log2(float(err) + 1) < LOGN + 20, (N1, np.log2(float(err) + 1))
        for l, mod in enumerate(params.qi[: levelQ + 1]):
            assert int(out[0][l].max()) < mod and int(out[1][l].max()) < mod          # canonical outputs
        assert np.log2(float(err) + 1) < LOGN + 8, (N1, np.log2(float(err) + 1))     # observed: ~2^9


def test_missing_galois_key_is_an_error():
    params, rng, s, sn, keys, vec, diag_coeffs, m, ct, levelQ, levelP = _setup(3, [1, 2], 0)
    del keys[params.GaloisElement(2)]
    ev = LT.Evaluator(params, keys)
    lt = LT.LinearTransformation(vec, levelQ, levelP, LOGN - 1, 0)
    out = [np.zeros((levelQ + 1, params.N()), dtype=U64) for _ in range(2)]
    try:
        ev.EvaluateMany(ct, [lt], [out])
    except KeyError as e:
        assert "GaloisKey" in str(e)
    else:
        raise AssertionError("expected a missing-key error")
