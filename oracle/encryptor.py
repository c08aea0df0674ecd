"""CPU restatement of the encryptor / key-generator inner loops -- SURVEY.md section 8(f) rank 3 -- with the samplers' OUTPUTS
as explicit arguments (the reference draws them from its blake2b-XOF PRNG, sampling/prng.go; what is restated here is
everything that happens to a sampled polynomial afterwards):

    Encryptor.encryptZeroPk            core/rlwe/encryptor.go:204-299   (Element[ring.Poly] case: one special prime, then ModDown)
    Encryptor.encryptZeroPkNoP         :301-341
    Encryptor.encryptZeroSkFromC1QP    :404-430   (and encryptZeroSk :346-395 with the uniform c1 given)
    KeyGenerator.genEvaluationKey      core/rlwe/keygenerator.go:287-330
    AddPolyTimesGadgetVectorToGadgetCiphertext   core/rlwe/gadgetciphertext.go:171-241

TEST INFRASTRUCTURE ONLY. Parity status: no bit-level vectors exist in the reference for these (its tests decrypt and bound the
noise); tests/test_oracle_encryptor.py pins the restatement the same way. Small polynomials are signed integer vectors; the
samplers' "negative zero" (a coefficient stored as q, ring/sampler_gaussian.go:178) only changes representatives that the
following NTT / CRed canonicalise, so it is not modelled."""
from __future__ import annotations

import numpy as np

from . import oracle as O

U64 = np.uint64


def small_rows(coeffs, moduli):
    """What a sampler's Read leaves in the rows of a polynomial: the signed value modulo each prime."""
    c = np.asarray(coeffs, dtype=np.int64)
    return np.stack([np.where(c < 0, c + np.int64(q), c).astype(U64) for q in moduli]) if len(moduli) else np.zeros((0, len(c)), dtype=U64)


class Encryptor:
    def __init__(self, params: O.Parameters):
        self.params = params
        self.be = O.BasisExtender(params.ringQ, params.ringP) if params.ringP is not None else None

    # --- core/rlwe/encryptor.go:204-299, ct = Element[ring.Poly] ---------------------------------------------------------
    def encryptZeroPk(self, levelQ, pk, u, e0, e1, is_ntt=True, is_montgomery=False):
        """pk = (pk0, pk1), each (nQ + nP, N) NTT + Montgomery rows (rlwe.PublicKey.Value). Returns (ct0, ct1)."""
        params = self.params
        levelP = 0                                                   # :221
        ringQ = params.ringQ.AtLevel(levelQ); ringP = params.ringP.AtLevel(levelP)
        N = params.N()
        nQk = len(params.qi)
        qs, ps = params.qi[: levelQ + 1], params.pi[: levelP + 1]

        def qp(small):                                               # sampler.Read + ExtendBasisSmallNormAndCenter
            return small_rows(small, qs), small_rows(small, ps)
        uQ, uP = qp(u)
        ringQ.NTT(uQ, uQ); ringP.NTT(uP, uP)
        ct = []
        for k, e in ((0, e0), (1, e1)):
            cQ = np.empty((levelQ + 1, N), dtype=U64); cP = np.empty((levelP + 1, N), dtype=U64)
            ringQ.MulCoeffsMontgomery(uQ, pk[k][: levelQ + 1], cQ); ringP.MulCoeffsMontgomery(uP, pk[k][nQk: nQk + levelP + 1], cP)
            ringQ.INTT(cQ, cQ); ringP.INTT(cP, cP)
            eQ, eP = qp(e)
            ringQ.Add(cQ, eQ, cQ); ringP.Add(cP, eP, cP)
            out = np.empty((levelQ + 1, N), dtype=U64)
            self.be.ModDownQPtoQ(levelQ, levelP, cQ, cP, out)
            if is_ntt:
                ringQ.NTT(out, out)
            if is_montgomery:
                ringQ.MForm(out, out)
            ct.append(out)
        return ct

    # --- :301-341 --------------------------------------------------------------------------------------------------------
    def encryptZeroPkNoP(self, levelQ, pk, u, e0, e1, is_ntt=True):
        ringQ = self.params.ringQ.AtLevel(levelQ)
        qs = self.params.qi[: levelQ + 1]
        N = self.params.N()
        uQ = small_rows(u, qs); ringQ.NTT(uQ, uQ)
        ct = []
        for k, e in ((0, e0), (1, e1)):
            c = np.empty((levelQ + 1, N), dtype=U64)
            ringQ.MulCoeffsMontgomery(uQ, pk[k][: levelQ + 1], c)
            eQ = small_rows(e, qs)
            if is_ntt:
                ringQ.NTT(eQ, eQ); ringQ.Add(c, eQ, c)
            else:
                ringQ.INTT(c, c); ringQ.Add(c, eQ, c)                # ReadAndAdd = CRed(a + b)
            ct.append(c)
        return ct

    # --- :404-430 (levelP = -1: Element[ring.Poly]) -------------------------------------------------------------------------
    def encryptZeroSkFromC1QP(self, levelQ, levelP, sk, c1, e, is_ntt=True, is_montgomery=False):
        """sk: (nQ + nP, N) NTT + Montgomery; c1: (levelQ+1 + levelP+1, N) uniform rows, NTT domain (modified in place when
        !is_ntt, like the reference). Returns c0 (QP-stacked rows)."""
        params = self.params
        ringQ = params.ringQ.AtLevel(levelQ)
        ringP = params.ringP.AtLevel(levelP) if levelP >= 0 else None
        nQk = len(params.qi)
        nq = levelQ + 1
        c0 = np.concatenate([small_rows(e, params.qi[:nq]), small_rows(e, params.pi[: levelP + 1])]) if levelP >= 0 else small_rows(e, params.qi[:nq])
        ringQ.NTT(c0[:nq], c0[:nq])
        if ringP is not None:
            ringP.NTT(c0[nq:], c0[nq:])
        if is_montgomery:
            ringQ.MForm(c0[:nq], c0[:nq])
            if ringP is not None:
                ringP.MForm(c0[nq:], c0[nq:])
        ringQ.MulCoeffsMontgomeryThenSub(c1[:nq], sk[:nq], c0[:nq])
        if ringP is not None:
            ringP.MulCoeffsMontgomeryThenSub(c1[nq:], sk[nQk: nQk + levelP + 1], c0[nq:])
        if not is_ntt:
            ringQ.INTT(c0[:nq], c0[:nq]); ringQ.INTT(c1[:nq], c1[:nq])
            if ringP is not None:
                ringP.INTT(c0[nq:], c0[nq:]); ringP.INTT(c1[nq:], c1[nq:])
        return c0


def add_poly_times_gadget_vector(params: O.Parameters, pt, evk: O.GadgetCiphertext):
    """AddPolyTimesGadgetVectorToGadgetCiphertext (core/rlwe/gadgetciphertext.go:171-241) for one GadgetCiphertext: adds
    pt * P * w^j to component 0 on the limbs of digit i. pt: (levelQ+1, N) NTT + Montgomery rows."""
    levelQ, levelP = evk.LevelQ(), evk.LevelP()
    ringQ = params.ringQ.AtLevel(levelQ)
    buff = np.empty((levelQ + 1, params.N()), dtype=U64)
    if levelP != -1:
        ringQ.MulScalarBigint(pt[: levelQ + 1], params.ringP.AtLevel(levelP).ModulusAtLevel[levelP], buff)
    else:
        levelP = 0
        buff[...] = pt[: levelQ + 1]
    sizes = evk.BaseTwoDecompositionVectorSize()
    for j in range(max(sizes)):
        for i in range(evk.data.shape[0]):
            if j < sizes[i]:
                for k in range(levelP + 1):
                    index = i * (levelP + 1) + k
                    if index >= levelQ + 1:
                        break
                    ringQ.SubRings[index].vecop("Add", evk.data[i, j, 0, index], buff[index], evk.data[i, j, 0, index])
        ringQ.MulScalar(buff, 1 << evk.BaseTwoDecomposition, buff)


def gen_evaluation_key(params: O.Parameters, skIn, skOut, a, e, pw2=0):
    """KeyGenerator.genEvaluationKey (core/rlwe/keygenerator.go:287-330) at the maximum levels with the samples given:
    a[i][j]: (nQ + nP, N) uniform rows (the key's component 1), e[i][j]: signed small polynomial. skIn: (nQ, N) rows, skOut:
    (nQ + nP, N) rows, both NTT + Montgomery."""
    levelQ, levelP = params.MaxLevelQ(), params.MaxLevelP()
    n = params.BaseRNSDecompositionVectorSize(levelQ, levelP)
    sizes = params.BaseTwoDecompositionVectorSize(levelQ, levelP, pw2)[:n]
    rows = levelQ + 1 + levelP + 1
    data = np.zeros((n, max(sizes), 2, rows, params.N()), dtype=U64)
    enc = Encryptor(params)
    for i in range(n):
        for j in range(sizes[i]):
            c1 = np.array(a[i][j], dtype=U64, copy=True)
            data[i, j, 0] = enc.encryptZeroSkFromC1QP(levelQ, levelP, skOut, c1, e[i][j], True, True)
            data[i, j, 1] = c1
    evk = O.GadgetCiphertext(data, levelQ + 1, levelP + 1, pw2, sizes)
    add_poly_times_gadget_vector(params, skIn, evk)
    return evk
