"""Pins the CPU oracle (oracle/) before anything trusts it.

1. Golden vectors: the reference's own NTT known-answer vectors
   (ring/ntt_test.go:10-89 -> tests/golden/ntt_vectors.json): NTT(poly) == polyNTT
   and INTT(NTT(x)) == x, exactly as TestNTT does (ring/ntt_test.go:91-119).
2. Big-integer property tests mirroring ring/ring_test.go: modular reduction edge operands
   (:537-673), MForm (:675-687), DivFloor/DivRound by last modulus (:245-334),
   ModUp/ModDown (:710-887), CI-NTT == standard 2N-NTT on symmetric input (:85-126).
"""
import json
import os
import random

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
U64 = np.uint64

Qi60 = [0x1fffffffffe00001, 0x1fffffffffc80001, 0x1fffffffffb40001, 0x1fffffffff500001,
        0x1fffffffff380001, 0x1fffffffff000001, 0x1ffffffffef00001, 0x1ffffffffee80001]
Pi60 = [0x1ffffffff6c80001, 0x1ffffffff6140001, 0x1ffffffff5f40001, 0x1ffffffff5700001]


def rand_poly(ring, rng):
    return np.stack([rng.integers(0, s.Modulus, ring.N(), dtype=U64) for s in ring.SubRings[: ring.level + 1]])


def test_golden_ntt_vectors():
    vecs = json.load(open(os.path.join(HERE, "golden", "ntt_vectors.json")))["vectors"]
    assert [v["N"] for v in vecs] == [16, 32, 64, 128, 256, 512]
    for v in vecs:
        ring = O.Ring(v["N"], v["Qis"])
        poly = np.array(v["poly"], dtype=U64)
        want = np.array(v["polyNTT"], dtype=U64)
        got = np.empty_like(poly)
        ring.NTT(poly, got)
        assert np.array_equal(got, want), f"NTT mismatch at N={v['N']}"
        # lazy output reduces to the same residues and stays below 6q (ring/ntt.go:66)
        lazy = np.empty_like(poly)
        ring.NTTLazy(poly, lazy)
        for i, q in enumerate(v["Qis"]):
            assert int(lazy[i].max()) < 6 * q
            assert np.array_equal(lazy[i] % U64(q), want[i])
        back = np.empty_like(poly)
        ring.INTT(got, back)
        assert np.array_equal(back, poly), f"INTT mismatch at N={v['N']}"
        back2 = np.empty_like(poly)
        ring.INTTLazy(got, back2)
        assert np.array_equal(back2, poly)


def test_tables_match_definition():
    """RootsForward[brev(j)] = psi^j * 2^64 mod q with psi = g^((q-1)/2N), g the smallest primitive root >= 3."""
    q, N = Qi60[0], 64
    s = O.SubRing(N, q)
    g = s.PrimitiveRoot
    assert g >= 3
    psi = pow(g, (q - 1) // (2 * N), q)
    assert pow(psi, N, q) == q - 1
    for j in (0, 1, 2, 5, 63):
        assert int(s.RootsForward[O.bit_reverse64(j, 6)]) == pow(psi, j, q) * (1 << 64) % q
        assert int(s.RootsBackward[O.bit_reverse64(j, 6)]) == pow(psi, -j, q) * (1 << 64) % q
    assert s.MRedConstant * q % (1 << 64) == 1
    assert s.NInv == pow(N, -1, q) * (1 << 64) % q


def test_modular_reduction_edge_operands():
    """ring/ring_test.go:537-673: BRed / MRed / BRedAdd on operands 1, q-1, 2^64-1 vs big-int."""
    L = O.lib()
    rnd = random.Random(1)
    for q in Qi60[:3] + [576460752303439873, 0x10001 * 0 + 35184372744193]:
        if not O.is_prime(q):
            continue
        brc = np.array(O.gen_bred_constant(q), dtype=U64)
        qinv = O.gen_mred_constant(q)
        R = 1 << 64
        xs = [0, 1, q - 1, q // 2] + [rnd.randrange(q) for _ in range(50)]
        for x in xs:
            for y in xs:
                assert L.lo_bred(x, y, q, brc.ctypes.data) == x * y % q
                assert L.lo_mred(x, y, q, qinv) == x * y * pow(R, -1, q) % q
                lz = L.lo_mred_lazy(x, y, q, qinv)
                assert lz < 2 * q and lz % q == x * y * pow(R, -1, q) % q
            assert L.lo_mform(x, q, brc.ctypes.data) == x * R % q
            assert L.lo_imform(x, q, qinv) == x * pow(R, -1, q) % q
        for a in [0, 1, q - 1, q, 2 * q + 1, (1 << 64) - 1] + [rnd.randrange(1 << 64) for _ in range(50)]:
            assert L.lo_bred_add(a, q, brc.ctypes.data) == a % q
            l2 = L.lo_bred_add_lazy(a, q, brc.ctypes.data)
            assert l2 < 2 * q and l2 % q == a % q
            assert L.lo_mform(a, q, brc.ctypes.data) == a * R % q   # MForm accepts any a < 2^64



def _div_round(x: int, d: int) -> int:
    """bignum.DivRound (utils/bignum/int.go): round(x/d), ties away from zero, sign aware."""
    q, r = divmod(abs(x), d)
    if 2 * r >= d:
        q += 1
    return q if x >= 0 else -q


@pytest.mark.parametrize("nb", [1, 2, 3])
def test_div_floor_round_by_last_modulus_many(nb):
    """ring/ring_test.go:245-334 (coefficient domain and NTT domain)."""
    rng = np.random.default_rng(7)
    N = 64
    ring = O.Ring(N, Qi60[:5])
    Q = ring.ModulusAtLevel[ring.level]
    rnd = random.Random(3)
    coeffs = [rnd.randrange(Q) for _ in range(N)]
    # centre like the test: values in (-Q/2, Q/2]
    cent = [c - Q if c > Q // 2 else c for c in coeffs]
    p0 = ring.NewPoly()
    ring.SetCoefficientsBigint(cent, p0)
    div = 1
    for i in range(nb):
        div *= ring.SubRings[ring.level - i].Modulus
    lvl_out = ring.level - nb
    rout = ring.AtLevel(lvl_out)
    Qout = ring.ModulusAtLevel[lvl_out]
    # floor
    buff, p1 = ring.NewPoly(), ring.NewPoly()
    ring.DivFloorByLastModulusMany(nb, p0.copy(), buff, p1)
    got = rout.PolyToBigint(p1[: lvl_out + 1])
    assert got == [(c // div) % Qout for c in cent]
    # round
    buff, p1 = ring.NewPoly(), ring.NewPoly()
    ring.DivRoundByLastModulusMany(nb, p0.copy(), buff, p1)
    got = rout.PolyToBigint(p1[: lvl_out + 1])
    assert got == [_div_round(c, div) % Qout for c in cent]
    # NTT-domain variants
    p0n = ring.NewPoly(); ring.NTT(p0, p0n)
    p1 = ring.NewPoly()
    ring.DivRoundByLastModulusManyNTT(nb, p0n.copy(), None, p1)
    back = rout.NewPoly(); rout.INTT(p1[: lvl_out + 1], back)
    assert rout.PolyToBigint(back) == [_div_round(c, div) % Qout for c in cent]
    p1 = ring.NewPoly()
    ring.DivFloorByLastModulusManyNTT(nb, p0n.copy(), p1)
    back = rout.NewPoly(); rout.INTT(p1[: lvl_out + 1], back)
    assert rout.PolyToBigint(back) == [(c // div) % Qout for c in cent]


def test_extend_basis_modup_moddown():
    """ring/ring_test.go:710-887: ModUpQtoP, ModUpPtoQ, ModDownQPtoQ, ModDownQPtoP vs big.Int."""
    N = 64
    ringQ = O.Ring(N, Qi60[:4]); ringP = O.Ring(N, Pi60[:3])
    be = O.BasisExtender(ringQ, ringP)
    rnd = random.Random(11)
    for levelQ in (0, 1, 3):
        for levelP in (0, 2):
            rq, rp = ringQ.AtLevel(levelQ), ringP.AtLevel(levelP)
            Q, P = ringQ.ModulusAtLevel[levelQ], ringP.ModulusAtLevel[levelP]
            # ModUp Q->P: centred representative of x mod Q, reduced mod P
            coeffs = [rnd.randrange(Q) for _ in range(N)]
            cent = [c - Q if c > Q // 2 else c for c in coeffs]   # test uses values in [-Q/2, Q/2]
            pq = rq.NewPoly(); rq.SetCoefficientsBigint(cent, pq)
            pp = rp.NewPoly()
            be.ModUpQtoP(levelQ, levelP, pq, pp)
            rp.Reduce(pp, pp)                                     # ring_test.go:745 reduces before comparing
            assert rp.PolyToBigint(pp) == [c % P for c in cent]
            # ModUp P->Q
            coeffs = [rnd.randrange(P) for _ in range(N)]
            cent = [c - P if c > P // 2 else c for c in coeffs]
            pp = rp.NewPoly(); rp.SetCoefficientsBigint(cent, pp)
            pq = rq.NewPoly()
            be.ModUpPtoQ(levelP, levelQ, pp, pq)
            rq.Reduce(pq, pq)
            assert rq.PolyToBigint(pq) == [c % Q for c in cent]
            # ModDown QP->Q : round(x / P) mod Q for x in (-QP/2, QP/2]
            QP = Q * P
            coeffs = [rnd.randrange(QP) for _ in range(N)]
            cent = [c - QP if c > QP // 2 else c for c in coeffs]
            pq = rq.NewPoly(); rq.SetCoefficientsBigint(cent, pq)
            pp = rp.NewPoly(); rp.SetCoefficientsBigint(cent, pp)
            out = rq.NewPoly()
            be.ModDownQPtoQ(levelQ, levelP, pq, pp, out)
            assert rq.PolyToBigint(out) == [_div_round(c, P) % Q for c in cent]
            # NTT variant
            pqn, ppn = rq.NewPoly(), rp.NewPoly()
            rq.NTT(pq, pqn); rp.NTT(pp, ppn)
            outn = rq.NewPoly()
            be.ModDownQPtoQNTT(levelQ, levelP, pqn, ppn, outn)
            back = rq.NewPoly(); rq.INTT(outn, back)
            assert rq.PolyToBigint(back) == [_div_round(c, P) % Q for c in cent]
            # ModDown QP->P : round(x / Q) mod P
            outp = rp.NewPoly()
            be.ModDownQPtoP(levelQ, levelP, pq, pp, outp)
            assert rp.PolyToBigint(outp) == [_div_round(c, Q) % P for c in cent]


def test_ntt_conjugate_invariant_matches_standard_2n():
    """ring/ring_test.go:85-126: CI-NTT of p equals the first half of the standard 2N-NTT of p unfolded
    symmetrically (p[0], p[1..N-1], 0, -p[N-1..1])."""
    N = 64
    q = Qi60[0]
    ci = O.SubRing(N, q, ring_type="ConjugateInvariant")
    std = O.SubRing(2 * N, q)
    assert ci.NthRoot == std.NthRoot == 4 * N
    rng = np.random.default_rng(5)
    p = rng.integers(0, q, N, dtype=U64)
    unfolded = np.zeros(2 * N, dtype=U64)
    unfolded[:N] = p
    for i in range(1, N):
        unfolded[2 * N - i] = (q - int(p[i])) % q
    want = np.empty(2 * N, dtype=U64); std.NTT(unfolded, want)
    got = np.empty(N, dtype=U64); ci.NTT(p, got)
    assert np.array_equal(got, want[:N])
    back = np.empty(N, dtype=U64); ci.INTT(got, back)
    assert np.array_equal(back, p)


def test_automorphism_ntt_index_is_galois_action():
    """AutomorphismNTT(NTT(p), g) == NTT(Automorphism(p, g)) (ring/automorphism.go)."""
    N = 128
    ring = O.Ring(N, Qi60[:2])
    rng = np.random.default_rng(9)
    p = rand_poly(ring, rng)
    for gal in (5, 25, 2 * N - 1, pow(5, 7, 2 * N)):
        a = ring.NewPoly(); ring.Automorphism(p, gal, a)
        an = ring.NewPoly(); ring.NTT(a, an)
        pn = ring.NewPoly(); ring.NTT(p, pn)
        bn = ring.NewPoly(); ring.AutomorphismNTT(pn, gal, bn)
        assert np.array_equal(an, bn)


def test_gen_moduli_shapes():
    q, p = O.gen_moduli(17, [56] + [45] * 5, [55, 55])
    assert len(set(q + p)) == 8
    for x, b in zip(q + p, [56] + [45] * 5 + [55, 55]):
        assert O.is_prime(x) and x % (1 << 17) == 1 and abs(np.log2(float(x)) - b) < 0.5
    # 61-bit primes are generated downstream only (core/rlwe/params.go:832-836): matches ring.Qi60
    q61, _ = O.gen_moduli(18, [61] * 4, [])
    assert q61 == Qi60[:4]
