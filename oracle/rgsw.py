"""CPU restatement of the RGSW external product -- SURVEY.md section 8(f) rank 4 -- `core/rgsw/evaluator.go` (reference):

    Evaluator.ExternalProduct                          :39-88     (dispatch on the RGSW ciphertext's levelP, ModDown)
    externalProduct32Bit                               :90-128    (levelQ = 0, no P, modulus below 2^29: lazy 64-bit accumulation)
    externalProductInPlaceSinglePAndBitDecomp          :130-208   (levelP < 1: RNS digit = one limb, optional base-2^w split)
    externalProductInPlaceMultipleP                    :210-283   (levelP >= 1: the gadget-product loop over both components)

TEST INFRASTRUCTURE ONLY, like the rest of oracle/. Parity status: the reference holds no bit-level vectors for this package
(core/rgsw/rgsw_test.go checks noise bounds); tests/test_oracle_rgsw.py pins the restatement by decryption on real RGSW
encryptions, and against the gadget-product restatement where the two coincide.

Conventions: an RLWE ciphertext is a pair of (levelQ+1, N) uint64 arrays in the NTT domain; an RGSW ciphertext is a pair of
oracle.GadgetCiphertext (rgsw.Ciphertext.Value, core/rgsw/elements.go:11-13)."""
from __future__ import annotations

import numpy as np

from . import oracle as O

U64 = np.uint64


class Evaluator:
    """rgsw.Evaluator over oracle.Evaluator."""

    def __init__(self, params: O.Parameters):
        self.params = params
        self.ev = O.Evaluator(params)

    # --- :39-88 -----------------------------------------------------------------------------------------------------
    def ExternalProduct(self, op0, op1, opOut):
        params = self.params
        levelQ, levelP = op1[0].LevelQ(), op1[0].LevelP()
        N = params.N()
        c0Q, c1Q = opOut[0][: levelQ + 1], opOut[1][: levelQ + 1]
        c0P = np.zeros((max(levelP + 1, 1), N), dtype=U64); c1P = np.zeros((max(levelP + 1, 1), N), dtype=U64)
        if op0[0] is opOut[0] or op0[1] is opOut[1]:                       # op0 == opOut: accumulate in buffers (:50-54)
            c0Q = np.zeros((levelQ + 1, N), dtype=U64); c1Q = np.zeros((levelQ + 1, N), dtype=U64)
        be = self.ev.BasisExtender
        if levelP < 1:
            ringQ = params.ringQ
            if levelQ == 0 and levelP == -1 and (ringQ.SubRings[0].Modulus >> 29) == 0:
                self.externalProduct32Bit(op0, op1, c0Q, c1Q)
                r0 = ringQ.AtLevel(0)
                r0.IMForm(c0Q, opOut[0][:1]); r0.IMForm(c1Q, opOut[1][:1])
            else:
                self.externalProductInPlaceSinglePAndBitDecomp(op0, op1, c0Q, c0P, c1Q, c1P)
                if levelP == 0:
                    be.ModDownQPtoQNTT(levelQ, levelP, c0Q, c0P, opOut[0][: levelQ + 1])
                    be.ModDownQPtoQNTT(levelQ, levelP, c1Q, c1P, opOut[1][: levelQ + 1])
                else:
                    opOut[0][: levelQ + 1] = c0Q; opOut[1][: levelQ + 1] = c1Q
        else:
            bQ = [np.zeros((levelQ + 1, N), dtype=U64) for _ in range(2)]; bP = [np.zeros((levelP + 1, N), dtype=U64) for _ in range(2)]
            self.externalProductInPlaceMultipleP(levelQ, levelP, op0, op1, bQ[0], bP[0], bQ[1], bP[1])
            # NB the reference ModDowns c0QP / c1QP, which alias the accumulators only when op0 == opOut; otherwise c0QP.Q is
            # opOut's own storage, which externalProductInPlaceMultipleP never wrote (:84-86). The accumulators are what is meant.
            be.ModDownQPtoQNTT(levelQ, levelP, bQ[0], bP[0], opOut[0][: levelQ + 1])
            be.ModDownQPtoQNTT(levelQ, levelP, bQ[1], bP[1], opOut[1][: levelQ + 1])

    # --- :90-128 ----------------------------------------------------------------------------------------------------
    def externalProduct32Bit(self, ct0, rgsw, c0, c1):
        ringQ = self.params.ringQ.AtLevel(0)
        s = ringQ.SubRings[0]
        N = self.params.N()
        pw2 = rgsw[0].BaseTwoDecomposition
        mask = (1 << pw2) - 1
        buffQ = np.empty((1, N), dtype=U64)
        cw = np.empty(N, dtype=U64); cwNTT = np.empty(N, dtype=U64)
        for i, el in enumerate(rgsw):
            ringQ.INTT(ct0[i][:1], buffQ)
            for j in range(el.BaseTwoDecompositionVectorSize()[0]):
                s.vecop("Mask", buffQ[0], None, cw, j * pw2, mask)
                s.NTTLazy(cw, cwNTT)
                name = "MulCoeffsLazy" if (i == 0 and j == 0) else "MulCoeffsLazyThenAddLazy"
                s.vecop(name, el.Q(0, j, 0)[0], cwNTT, c0[0])
                s.vecop(name, el.Q(0, j, 1)[0], cwNTT, c1[0])

    # --- :130-208 ---------------------------------------------------------------------------------------------------
    def externalProductInPlaceSinglePAndBitDecomp(self, ct0, rgsw, c0Q, c0P, c1Q, c1P):
        params = self.params
        levelQ, levelP = rgsw[0].LevelQ(), rgsw[0].LevelP()
        ringQ = params.ringQ.AtLevel(levelQ)
        ringP = params.ringP.AtLevel(levelP) if levelP >= 0 else None
        N = params.N()
        pw2 = rgsw[0].BaseTwoDecomposition
        mask = (1 << pw2) - 1
        if mask == 0:
            mask = 0xFFFFFFFFFFFFFFFF
        nRNS = rgsw[0].data.shape[0]                       # BaseRNSDecompositionVectorSize() = len(Value)
        nPw2 = rgsw[0].BaseTwoDecompositionVectorSize()
        buffQ = np.empty((levelQ + 1, N), dtype=U64)
        cw = np.empty(N, dtype=U64); cwNTT = np.empty(N, dtype=U64)
        for k, el in enumerate(rgsw):
            ringQ.INTT(ct0[k][: levelQ + 1], buffQ)
            for i in range(nRNS):
                for j in range(nPw2[i]):
                    s0 = ringQ.SubRings[0]
                    s0.vecop("Mask", buffQ[i], None, cw, j * pw2, mask)      # ring.MaskVec is modulus-free
                    name = "MulCoeffsMontgomery" if (k == 0 and i == 0 and j == 0) else "MulCoeffsMontgomeryThenAdd"
                    for u, s in enumerate(ringQ.SubRings[: levelQ + 1]):
                        s.NTTLazy(cw, cwNTT)
                        s.vecop(name, el.Q(i, j, 0)[u], cwNTT, c0Q[u])
                        s.vecop(name, el.Q(i, j, 1)[u], cwNTT, c1Q[u])
                    if ringP is not None:
                        for u, s in enumerate(ringP.SubRings[: levelP + 1]):
                            s.NTTLazy(cw, cwNTT)
                            s.vecop(name, el.P(i, j, 0)[u], cwNTT, c0P[u])
                            s.vecop(name, el.P(i, j, 1)[u], cwNTT, c1P[u])

    # --- :210-283 ---------------------------------------------------------------------------------------------------
    def externalProductInPlaceMultipleP(self, levelQ, levelP, ct0, rgsw, c0Q, c0P, c1Q, c1P):
        params = self.params
        ringQ = params.ringQ.AtLevel(levelQ); ringP = params.ringP.AtLevel(levelP)
        N = params.N()
        n = params.BaseRNSDecompositionVectorSize(levelQ, levelP)
        QiOverF = params.QiOverflowMargin(levelQ) >> 1
        PiOverF = params.PiOverflowMargin(levelP) >> 1
        c2Q = np.empty((levelQ + 1, N), dtype=U64); c2P = np.empty((levelP + 1, N), dtype=U64)
        buffQ = np.empty((levelQ + 1, N), dtype=U64)
        reduce = 0
        for k, el in enumerate(rgsw):
            c2NTT = ct0[k][: levelQ + 1]
            ringQ.INTT(c2NTT, buffQ)
            for i in range(n):
                self.ev.DecomposeSingleNTT(levelQ, levelP, levelP + 1, i, c2NTT, buffQ, c2Q, c2P)
                name = "MulCoeffsMontgomeryLazy" if (k == 0 and i == 0) else "MulCoeffsMontgomeryLazyThenAddLazy"
                getattr(ringQ, name)(el.Q(i, 0, 0)[: levelQ + 1], c2Q, c0Q); getattr(ringP, name)(el.P(i, 0, 0)[: levelP + 1], c2P, c0P)
                getattr(ringQ, name)(el.Q(i, 0, 1)[: levelQ + 1], c2Q, c1Q); getattr(ringP, name)(el.P(i, 0, 1)[: levelP + 1], c2P, c1P)
                if reduce % QiOverF == QiOverF - 1:
                    ringQ.Reduce(c0Q, c0Q); ringQ.Reduce(c1Q, c1Q)
                if reduce % PiOverF == PiOverF - 1:
                    ringP.Reduce(c0P, c0P); ringP.Reduce(c1P, c1P)
                reduce += 1
        if reduce % QiOverF != 0:
            ringQ.Reduce(c0Q, c0Q); ringQ.Reduce(c1Q, c1Q)
        if reduce % PiOverF != 0:
            ringP.Reduce(c0P, c0P); ringP.Reduce(c1P, c1P)
