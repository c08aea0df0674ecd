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
