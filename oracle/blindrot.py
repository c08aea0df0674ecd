"""CPU restatement of blindrot.Evaluator.BlindRotateCore -- `core/rgsw/blindrot/evaluator.go:144-283` (Algorithm 3 of eprint 2022/198):

    BlindRotateCore                    :144-203
    evaluateFromDiscreteLogSets        :206-229
    getGaloisElementInverseMap         :232-258
    getDiscreteLogSets                 :261-283

TEST INFRASTRUCTURE ONLY. Built from the pinned restatements of Evaluator.Automorphism (oracle/oracle.py) and rgsw ExternalProduct
(oracle/rgsw.py); it adds only control flow, which tests/test_oracle_blindrot.py pins by decryption: with RGSW(X^{s_j}) keys the
accumulator f(X) ends up multiplied by X^{<a, s>} (up to noise)."""
from __future__ import annotations

from typing import Dict, List

import numpy as np

from . import oracle as O
from . import rgsw as RG

WINDOW_SIZE = 10          # core/rgsw/blindrot/keys.go:14
GALOIS_GEN = 5


def galois_element_inverse_map(N: int) -> Dict[int, int]:
    twoN, nhalf = N << 1, N >> 1
    out, pw = {}, 1
    for i in range(nhalf):
        out[pw] = i
        out[twoN - pw] = -i
        pw = pw * GALOIS_GEN & (twoN - 1)
    return out


def discrete_log_sets(a, N: int) -> Dict[int, List[int]]:
    dl = galois_element_inverse_map(N)
    sets: Dict[int, List[int]] = {}
    for i, ai in enumerate(a):
        ai = int(ai)
        if ai & 1 != 1 and ai != 0:
            raise ValueError("getDiscreteLogSets: a[i] is not odd")
        sets.setdefault(dl.get(ai, 0), []).append(i)          # a missing key reads as Go's zero value
    return sets


class Evaluator:
    def __init__(self, params: O.Parameters, window_size: int = WINDOW_SIZE):
        self.params = params
        self.ev = O.Evaluator(params)
        self.rg = RG.Evaluator(params)
        self.window = window_size

    def BlindRotateCore(self, a, acc, brk: List[list], galois_keys: Dict[int, O.GadgetCiphertext], level=None):
        """acc: [c0, c1] NTT-domain rows (modified in place); brk[j] = [Value0, Value1] of RGSW(X^{s_j})."""
        params = self.params
        N = params.N()
        level = len(acc[0]) - 1 if level is None else level
        sets = discrete_log_sets(a, N)

        def automorph(k):
            g = params.GaloisElement(k)
            if g not in galois_keys:
                raise KeyError("GaloisKey[%d] is missing" % g)
            out = [np.empty_like(acc[0]), np.empty_like(acc[1])]
            self.ev.Automorphism([acc[0], acc[1]], g, galois_keys[g], out, level)
            acc[0][...] = out[0]; acc[1][...] = out[1]

        def from_sets(k, v):
            if k in sets:
                if v != 0:
                    automorph(v); v = 0
                for j in sets[k]:
                    self.rg.ExternalProduct(acc, brk[j], acc)
            v += 1
            if v == self.window or k == 1:
                automorph(v); v = 0
            return v
        v = 0
        nhalf = N >> 1
        for i in range(nhalf - 1, 0, -1):
            v = from_sets(-i, v)
        from_sets(N << 1, 0)
        g = 2 * N - GALOIS_GEN
        if g not in galois_keys:
            raise KeyError("GaloisKey[%d] is missing" % g)
        out = [np.empty_like(acc[0]), np.empty_like(acc[1])]
        self.ev.Automorphism([acc[0], acc[1]], g, galois_keys[g], out, level)
        acc[0][...] = out[0]; acc[1][...] = out[1]
        for i in range(nhalf - 1, 0, -1):
            v = from_sets(i, v)
        from_sets(0, 0)
