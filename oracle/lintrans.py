"""CPU restatement of the hoisted linear-transformation evaluator — SURVEY.md section 8(f) rank 1, the caller right above
the key-switch path — `circuits/common/lintrans/lintrans_evaluator.go` (reference, pure Go):

    EvaluateMany                                           :41-79    (decompose once, pre-rotate, dispatch)
    PreRotatedCiphertextForDiagonalMatrixMultiplication    :82-110   (AutomorphismHoistedLazy per baby step)
    MultiplyByDiagMatrix                                   :141-274  (single hoisting, one key per diagonal)
    MultiplyByDiagMatrixBSGS                               :280-470  (double hoisting, baby-step giant-step)
    BSGSIndex                                              circuits/common/lintrans/lintrans.go:344-367

TEST INFRASTRUCTURE ONLY, like the rest of oracle/: it is built from the oracle's Ring / Evaluator restatements
(oracle/oracle.py) and is what a device implementation of this row will be compared with. Parity status: the reference
holds no bit-level vectors for this evaluator; tests/test_oracle_lintrans.py pins it by decrypt-style algebraic checks
on real keys (result == sum_k diag_k * rot_k(message) up to key-switch noise) and by BSGS == naive agreement.

Conventions: ciphertexts are pairs of (levelQ+1, N) uint64 arrays in the NTT domain; a diagonal is a pair (Q rows, P rows)
in NTT + Montgomery form (what LinearTransformation.Vec holds, lintrans.go:150-236); Galois keys are
{galEl: oracle.GadgetCiphertext}."""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np

from . import oracle as O

U64 = np.uint64


def bsgs_index(non_zero_diags: Sequence[int], slots: int, N1: int):
    """BSGSIndex, lintrans.go:344-367: {giant step: sorted baby steps}, sorted giant steps, sorted baby steps."""
    index: Dict[int, List[int]] = {}
    rotN1, rotN2 = set(), set()
    for rot in non_zero_diags:
        rot &= slots - 1
        idxN1 = ((rot // N1) * N1) & (slots - 1)
        idxN2 = rot & (N1 - 1)
        index.setdefault(idxN1, []).append(idxN2)
        rotN1.add(idxN1); rotN2.add(idxN2)
    for k in index:
        index[k].sort()
    return index, sorted(rotN1), sorted(rotN2)


class LinearTransformation:
    """lintrans.LinearTransformation (lintrans.go:150-160): Vec[k] = (Q rows, P rows) NTT + Montgomery, N1 = 0 selects the
    naive evaluator."""

    def __init__(self, vec: Dict[int, Tuple[np.ndarray, np.ndarray]], levelQ: int, levelP: int, log_slots: int, N1: int = 0):
        self.Vec, self.LevelQ, self.LevelP, self.LogSlots, self.N1 = vec, levelQ, levelP, log_slots, N1

    def BSGSIndex(self):
        return bsgs_index(list(self.Vec.keys()), 1 << self.LogSlots, self.N1)


class Evaluator:
    """lintrans.Evaluator over oracle.Evaluator (the embedded schemes evaluator only contributes GetRLWEParameters,
    CheckAndGetGaloisKey and the buffers)."""

    def __init__(self, params: O.Parameters, galois_keys: Dict[int, O.GadgetCiphertext]):
        self.params = params
        self.ev = O.Evaluator(params)
        self.keys = galois_keys

    def _key(self, galEl: int) -> O.GadgetCiphertext:
        if galEl not in self.keys:
            raise KeyError("GaloisKey[%d] is missing" % galEl)       # CheckAndGetGaloisKey
        return self.keys[galEl]

    # --- :82-110 ---------------------------------------------------------------------------------------------
    def PreRotatedCiphertextForDiagonalMatrixMultiplication(self, levelQ, levelP, ctIn, decompQ, decompP, rots, ctPreRot: dict):
        for i in list(ctPreRot.keys()):
            if i not in rots:
                del ctPreRot[i]
        N = self.params.N()
        for i in rots:
            if i != 0 and i not in ctPreRot:
                q = [np.zeros((levelQ + 1, N), dtype=U64) for _ in range(2)]
                p = [np.zeros((levelP + 1, N), dtype=U64) for _ in range(2)]
                galEl = self.params.GaloisElement(i)
                self.ev.AutomorphismHoistedLazy(levelQ, ctIn, decompQ, decompP, galEl, self._key(galEl), q, p)
                ctPreRot[i] = (q, p)

    # --- :41-79 ----------------------------------------------------------------------------------------------
    def EvaluateMany(self, ctIn, matrices: Sequence[LinearTransformation], opOut: Sequence[list]):
        levelQ = min(len(ctIn[0]) - 1, max(m.LevelQ for m in matrices))      # levelQ of the decomposition (:47-50)
        levelP = matrices[0].LevelP
        N = self.params.N()
        n = self.params.BaseRNSDecompositionVectorSize(levelQ, levelP)
        decompQ = [np.zeros((levelQ + 1, N), dtype=U64) for _ in range(n)]
        decompP = [np.zeros((levelP + 1, N), dtype=U64) for _ in range(n)]
        self.ev.DecomposeNTT(levelQ, levelP, levelP + 1, ctIn[1][: levelQ + 1].copy(), True, decompQ, decompP)
        ctPreRot: dict = {}
        for m, out in zip(matrices, opOut):
            if m.N1 == 0:
                self.MultiplyByDiagMatrix(ctIn, m, decompQ, decompP, out)
            else:
                _, _, rotN2 = m.BSGSIndex()
                self.PreRotatedCiphertextForDiagonalMatrixMultiplication(levelQ, levelP, ctIn, decompQ, decompP, rotN2, ctPreRot)
                self.MultiplyByDiagMatrixBSGS(ctIn, m, ctPreRot, out)

    # --- :141-274 --------------------------------------------------------------------------------------------
    def MultiplyByDiagMatrix(self, ctIn, matrix: LinearTransformation, decompQ, decompP, opOut):
        params = self.params
        levelQ = min(len(opOut[0]) - 1, len(ctIn[0]) - 1, matrix.LevelQ)
        levelP = matrix.LevelP
        ringQ = params.ringQ.AtLevel(levelQ); ringP = params.ringP.AtLevel(levelP)
        N = params.N()
        QiOverF = params.QiOverflowMargin(levelQ); PiOverF = params.PiOverflowMargin(levelP)
        nq, npp = levelQ + 1, levelP + 1
        c0Q, c1Q = opOut[0][:nq], opOut[1][:nq]
        c0P = np.zeros((npp, N), dtype=U64); c1P = np.zeros((npp, N), dtype=U64)
        ct0 = ctIn[0][:nq].copy(); ct1 = ctIn[1][:nq].copy()
        ct0TimesP = np.zeros((nq, N), dtype=U64)
        ringQ.MulScalarBigint(ct0, ringP.ModulusAtLevel[levelP], ct0TimesP)
        slots = 1 << matrix.LogSlots
        keys = sorted(matrix.Vec.keys())
        state = False
        if keys and keys[0] == 0:
            state = True
            keys = keys[1:]
        tQ = [np.zeros((nq, N), dtype=U64) for _ in range(2)]; tP = [np.zeros((npp, N), dtype=U64) for _ in range(2)]
        aQ = [np.zeros((nq, N), dtype=U64) for _ in range(2)]; aP = [np.zeros((npp, N), dtype=U64) for _ in range(2)]
        for i, k0 in enumerate(keys):
            k = k0 & (slots - 1)
            galEl = params.GaloisElement(k)
            evk = self._key(galEl)
            assert evk.LevelP() == levelP, "LinearTransformation.LevelP != GaloisKey.LevelP()"
            index = ringQ.AutomorphismNTTIndex(galEl)
            self.ev.GadgetProductHoistedLazy(levelQ, decompQ, decompP, evk, aQ, aP)
            ringQ.Add(aQ[0], ct0TimesP, aQ[0])
            ringQ.AutomorphismNTTWithIndex(aQ[0], index, tQ[0]); ringP.AutomorphismNTTWithIndex(aP[0], index, tP[0])
            ringQ.AutomorphismNTTWithIndex(aQ[1], index, tQ[1]); ringP.AutomorphismNTTWithIndex(aP[1], index, tP[1])
            ptQ, ptP = matrix.Vec[k0]                # NB the reference indexes Vec with the masked k; keys are < slots there
            op = "MulCoeffsMontgomery" if i == 0 else "MulCoeffsMontgomeryThenAdd"
            getattr(ringQ, op)(ptQ[:nq], tQ[0], c0Q); getattr(ringP, op)(ptP[:npp], tP[0], c0P)
            getattr(ringQ, op)(ptQ[:nq], tQ[1], c1Q); getattr(ringP, op)(ptP[:npp], tP[1], c1P)
            if i % QiOverF == QiOverF - 1:
                ringQ.Reduce(c0Q, c0Q); ringQ.Reduce(c1Q, c1Q)
            if i % PiOverF == PiOverF - 1:
                ringP.Reduce(c0P, c0P); ringP.Reduce(c1P, c1P)
        if len(keys) % QiOverF == 0:
            ringQ.Reduce(c0Q, c0Q); ringQ.Reduce(c1Q, c1Q)
        if len(keys) % PiOverF == 0:
            ringP.Reduce(c0P, c0P); ringP.Reduce(c1P, c1P)
        be = self.ev.BasisExtender
        be.ModDownQPtoQNTT(levelQ, levelP, c0Q, c0P, c0Q)
        be.ModDownQPtoQNTT(levelQ, levelP, c1Q, c1P, c1Q)
        if state:
            ringQ.MulCoeffsMontgomeryThenAdd(matrix.Vec[0][0][:nq], ct0, c0Q)
            ringQ.MulCoeffsMontgomeryThenAdd(matrix.Vec[0][0][:nq], ct1, c1Q)

    # --- :280-470 --------------------------------------------------------------------------------------------
    def MultiplyByDiagMatrixBSGS(self, ctIn, matrix: LinearTransformation, ctInPreRot: dict, opOut):
        params = self.params
        levelQ = min(len(opOut[0]) - 1, len(ctIn[0]) - 1, matrix.LevelQ)
        levelP = matrix.LevelP
        ringQ = params.ringQ.AtLevel(levelQ); ringP = params.ringP.AtLevel(levelP)
        N = params.N()
        nq, npp = levelQ + 1, levelP + 1
        QiOverF = params.QiOverflowMargin(levelQ) >> 1; PiOverF = params.PiOverflowMargin(levelP) >> 1
        index, _, _ = matrix.BSGSIndex()
        ct0 = ctIn[0][:nq].copy(); ct1 = ctIn[1][:nq].copy()
        t0Q = np.zeros((nq, N), dtype=U64); t0P = np.zeros((npp, N), dtype=U64)      # inner-loop accumulators
        t1Q = np.zeros((nq, N), dtype=U64); t1P = np.zeros((npp, N), dtype=U64)
        cQ = [np.zeros((nq, N), dtype=U64) for _ in range(2)]; cP = [np.zeros((npp, N), dtype=U64) for _ in range(2)]
        c0Q, c1Q = opOut[0][:nq], opOut[1][:nq]
        c0P = np.zeros((npp, N), dtype=U64); c1P = np.zeros((npp, N), dtype=U64)
        Pm = ringP.ModulusAtLevel[levelP]
        ringQ.MulScalarBigint(ct0, Pm, ct0)
        ringQ.MulScalarBigint(ct1, Pm, ct1)
        be = self.ev.BasisExtender
        cnt0 = 0
        for j in sorted(index.keys()):
            cnt1 = 0
            for i in index[j]:
                ptQ, ptP = matrix.Vec[j + i]
                ptQ = ptQ[:nq]; ptP = ptP[:npp]
                if i == 0:
                    if cnt1 == 0:
                        ringQ.MulCoeffsMontgomeryLazy(ptQ, ct0, t0Q); ringQ.MulCoeffsMontgomeryLazy(ptQ, ct1, t1Q)
                        t0P[...] = 0; t1P[...] = 0
                    else:
                        ringQ.MulCoeffsMontgomeryLazyThenAddLazy(ptQ, ct0, t0Q); ringQ.MulCoeffsMontgomeryLazyThenAddLazy(ptQ, ct1, t1Q)
                else:
                    (rq, rp) = ctInPreRot[i]
                    op = "MulCoeffsMontgomeryLazy" if cnt1 == 0 else "MulCoeffsMontgomeryLazyThenAddLazy"
                    getattr(ringQ, op)(ptQ, rq[0], t0Q); getattr(ringP, op)(ptP, rp[0], t0P)
                    getattr(ringQ, op)(ptQ, rq[1], t1Q); getattr(ringP, op)(ptP, rp[1], t1P)
                if cnt1 % QiOverF == QiOverF - 1:
                    ringQ.Reduce(t0Q, t0Q); ringQ.Reduce(t1Q, t1Q)
                if cnt1 % PiOverF == PiOverF - 1:
                    ringP.Reduce(t0P, t0P); ringP.Reduce(t1P, t1P)
                cnt1 += 1
            if cnt1 % QiOverF != 0:
                ringQ.Reduce(t0Q, t0Q); ringQ.Reduce(t1Q, t1Q)
            if cnt1 % PiOverF != 0:
                ringP.Reduce(t0P, t0P); ringP.Reduce(t1P, t1P)
            if j != 0:
                be.ModDownQPtoQNTT(levelQ, levelP, t1Q, t1P, t1Q)              # hoisted ModDown of the c1 part
                galEl = params.GaloisElement(j)
                evk = self._key(galEl)
                assert evk.LevelP() == levelP
                rot = ringQ.AutomorphismNTTIndex(galEl)
                self.ev.GadgetProductLazy(levelQ, t1Q.copy(), evk, cQ, cP)
                ringQ.Add(cQ[0], t0Q, cQ[0]); ringP.Add(cP[0], t0P, cP[0])
                if cnt0 == 0:
                    ringQ.AutomorphismNTTWithIndex(cQ[0], rot, c0Q); ringP.AutomorphismNTTWithIndex(cP[0], rot, c0P)
                    ringQ.AutomorphismNTTWithIndex(cQ[1], rot, c1Q); ringP.AutomorphismNTTWithIndex(cP[1], rot, c1P)
                else:
                    ringQ.AutomorphismNTTWithIndexThenAddLazy(cQ[0], rot, c0Q); ringP.AutomorphismNTTWithIndexThenAddLazy(cP[0], rot, c0P)
                    ringQ.AutomorphismNTTWithIndexThenAddLazy(cQ[1], rot, c1Q); ringP.AutomorphismNTTWithIndexThenAddLazy(cP[1], rot, c1P)
            else:
                if cnt0 == 0:
                    c0Q[...] = t0Q; c0P[...] = t0P; c1Q[...] = t1Q; c1P[...] = t1P
                else:
                    ringQ.AddLazy(c0Q, t0Q, c0Q); ringP.AddLazy(c0P, t0P, c0P)
                    ringQ.AddLazy(c1Q, t1Q, c1Q); ringP.AddLazy(c1P, t1P, c1P)
            if cnt0 % QiOverF == QiOverF - 1:
                ringQ.Reduce(c0Q, c0Q); ringQ.Reduce(c1Q, c1Q)
            if cnt0 % PiOverF == PiOverF - 1:
                ringP.Reduce(c0P, c0P); ringP.Reduce(c1P, c1P)
            cnt0 += 1
        if cnt0 % QiOverF != 0:
            ringQ.Reduce(c0Q, c0Q); ringQ.Reduce(c1Q, c1Q)
        if cnt0 % PiOverF != 0:
            ringP.Reduce(c0P, c0P); ringP.Reduce(c1P, c1P)
        be.ModDownQPtoQNTT(levelQ, levelP, c0Q, c0P, c0Q)
        be.ModDownQPtoQNTT(levelQ, levelP, c1Q, c1P, c1Q)
