"""The native CPU batch loop of bench.py's reference arm (oracle/lattigo_cpu_batch.c) is bit-identical to the oracle's
CKKSEvaluator (oracle/oracle.py) -- so the timed CPU baseline computes exactly what the parity tests check -- and its
specialised transform loop nests reproduce the reference's golden NTT vectors (ring/ntt_test.go:10-89)."""
import json
import os

import numpy as np
import pytest

from oracle import cpu_batch as CB
from oracle import oracle as O
from tests import helpers as H

U64 = np.uint64
HERE = os.path.dirname(os.path.abspath(__file__))


def test_fast_transforms_on_golden_vectors():
    vecs = json.load(open(os.path.join(HERE, "golden", "ntt_vectors.json")))["vectors"]
    L = CB.lib()
    for v in vecs:
        N = v["N"]
        poly = np.array(v["poly"], dtype=U64); want = np.array(v["polyNTT"], dtype=U64)
        for i, q in enumerate(v["Qis"]):
            s = O.get_subring(N, q)
            out = np.empty(N, dtype=U64)
            L.lb_ntt(poly[i].ctypes.data, out.ctypes.data, N, q, s.MRedConstant, s.brc.ctypes.data, s.RootsForward.ctypes.data)
            assert np.array_equal(out, want[i])
            lazy = np.empty(N, dtype=U64); lazy_o = np.empty(N, dtype=U64)
            L.lb_ntt_lazy(poly[i].ctypes.data, lazy.ctypes.data, N, q, s.MRedConstant, s.RootsForward.ctypes.data)
            s.NTTLazy(np.ascontiguousarray(poly[i]), lazy_o)
            assert np.array_equal(lazy, lazy_o)          # the exact [0, 6q) representative, not only the residue
            back = np.empty(N, dtype=U64)
            L.lb_intt(out.ctypes.data, back.ctypes.data, N, s.NInv, q, s.MRedConstant, s.RootsBackward.ctypes.data)
            assert np.array_equal(back, poly[i])


@pytest.mark.parametrize("case", ["k2_ragged", "k3_single_limb", "k4_full", "q61"])
def test_batch_matches_oracle(case):
    logN = 8
    if case == "k2_ragged":
        q, p = O.gen_moduli(logN + 1, [56, 45, 45, 45, 45], [55, 55]); levels = (4, 3, 2)
    elif case == "k3_single_limb":
        q, p = O.gen_moduli(logN + 1, [56, 45, 45, 45, 45, 45, 45], [55, 55, 55]); levels = (6, 5, 4, 3)
    elif case == "k4_full":
        q, p = O.gen_moduli(logN + 1, [56] + [45] * 7, [55] * 4); levels = (7, 5, 4)
    else:
        q, p = H.Qi60[:5], H.Pi60[:2]; levels = (4, 2)
    params = O.Parameters(logN, q, p)
    N = params.N()
    rng = np.random.default_rng(99)
    rlk = H.random_gadget_ciphertext(params, params.MaxLevelQ(), params.MaxLevelP(), rng)
    ev = O.CKKSEvaluator(params, rlk)
    for level in levels:
        npairs = 3
        a = np.stack([np.stack([H.rand_poly(q[: level + 1], N, rng) for _ in range(2)]) for _ in range(npairs)])
        b = np.stack([np.stack([H.rand_poly(q[: level + 1], N, rng) for _ in range(2)]) for _ in range(npairs)])
        plan = CB.CKKSBatchPlan(params, rlk, level)
        _, per, out, mid = plan.run(a, b, npairs, nthreads=2, store=True, store_mid=True)
        assert (per > 0).all()
        for i in range(npairs):
            m = ev.MulRelinNew([a[i, 0], a[i, 1]], [b[i, 0], b[i, 1]])
            r = ev.Rescale(m)
            assert np.array_equal(mid[i, 0], m[0]) and np.array_equal(mid[i, 1], m[1]), (case, level, i)
            assert np.array_equal(out[i, 0], r[0]) and np.array_equal(out[i, 1], r[1]), (case, level, i)
        # shared-input mode (RunParallel-style): every pair reads pair 0
        _, _, out1, _ = plan.run(a[:1].copy(), b[:1].copy(), 2, nthreads=2, store=True)
        assert np.array_equal(out1[0], out[0]) and np.array_equal(out1[1], out[0])
    CB.lib().lo_batch_release()


def test_usable_cpus_reports_something():
    u = CB.usable_cpus()
    assert 1 <= u["usable"] <= u["affinity"]
