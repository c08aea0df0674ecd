"""Host-side mirror of core/rgsw (rgsw.Ciphertext, rgsw.Evaluator.ExternalProduct) bound to the C ABI."""
from __future__ import annotations

from . import _lib
from .ring import Context, _dptr, _stream
from .rlwe import GadgetCiphertext


class Ciphertext:
    """rgsw.Ciphertext (core/rgsw/elements.go:11-13): Value = [2]rlwe.GadgetCiphertext, device resident."""

    def __init__(self, v0: GadgetCiphertext, v1: GadgetCiphertext):
        self.Value = [v0, v1]

    def LevelQ(self): return self.Value[0].LevelQ()
    def LevelP(self): return self.Value[0].LevelP()


class Evaluator:
    """rgsw.Evaluator (core/rgsw/evaluator.go:12-24)."""

    def __init__(self, ctx: Context):
        self.ctx = ctx

    def ExternalProduct(self, op0, op1: Ciphertext, opOut):
        """op0 / opOut: (batch, 2, level+1, N) or (2, level+1, N) NTT-domain RLWE ciphertexts; rows 0..op1.LevelQ() of opOut are written."""
        batch = op0.shape[0] if op0.dim() == 4 else 1
        _lib.check(_lib.lib().lgpu_rgsw_external_product(self.ctx.h, _dptr(op0), op0.shape[-2] - 1, op1.Value[0].ref(), op1.Value[1].ref(),
                                                         _dptr(opOut), opOut.shape[-2] - 1, batch, _stream()))


class BlindRotationEvaluator:
    """blindrot.Evaluator.BlindRotateCore (core/rgsw/blindrot/evaluator.go:144-229) on the device: the accumulator-loop of Algorithm 3
    (eprint 2022/198), driven by the host-side LWE mask `a` (values modulo 2N), in place on `acc` ((2, level+1, N), NTT domain)."""

    def __init__(self, ctx: Context, blind_rotation_keys, automorphism_keys, window_size: int = 10):
        import ctypes
        self.ctx = ctx
        self.brk = list(blind_rotation_keys)                      # rgsw.Ciphertext per LWE coefficient: RGSW(X^{s_j})
        n = len(self.brk)
        self._b0 = (_lib.GadgetCtStruct * n)(*[k.Value[0].struct for k in self.brk])
        self._b1 = (_lib.GadgetCtStruct * n)(*[k.Value[1].struct for k in self.brk])
        self.keys = dict(automorphism_keys)                       # {galEl: GadgetCiphertext}
        els = list(self.keys.keys())
        self._els = (ctypes.c_uint64 * max(1, len(els)))(*els)
        self._structs = (_lib.GadgetCtStruct * max(1, len(els)))(*[self.keys[g].struct for g in els])
        self._gks = _lib.GaloisKeysStruct(len(els), self._els, self._structs)
        self.window = window_size

    def BlindRotateCore(self, a, acc):
        import ctypes
        import numpy as np
        a = np.ascontiguousarray(np.asarray(a, dtype=np.uint64))
        assert len(a) == len(self.brk)
        _lib.check(_lib.lib().lgpu_blind_rotate_core(self.ctx.h, a.ctypes.data, len(a), _dptr(acc), acc.shape[-2] - 1, self._b0, self._b1,
                                                     ctypes.byref(self._gks), self.window, _stream()))
