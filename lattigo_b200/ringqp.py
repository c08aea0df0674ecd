"""Host-side mirror of ring/ringqp.Ring (ring/ringqp/ring.go:15-49, operations.go:8-316): every operation is the
RingQ operation on the .Q part followed by the RingP operation on the .P part. A ringqp.Poly is a pair of device
polynomials (Q rows, P rows); P may be None (levelP = -1)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

from .ring import Context, Ring, OPS


@dataclass
class Poly:
    """ringqp.Poly (ring/ringqp/poly.go): Q and P parts."""
    Q: object
    P: Optional[object] = None


class RingQP:
    def __init__(self, ctx: Context, levelQ: Optional[int] = None, levelP: Optional[int] = None):
        self.ctx = ctx
        self.RingQ: Ring = ctx.ringQ if levelQ is None else ctx.ringQ.AtLevel(levelQ)
        self.RingP: Optional[Ring] = None
        if ctx.ringP is not None and (levelP is None or levelP >= 0):
            self.RingP = ctx.ringP if levelP is None else ctx.ringP.AtLevel(levelP)

    def AtLevel(self, levelQ: int, levelP: int) -> "RingQP":      # ring/ringqp/ring.go:33-49
        return RingQP(self.ctx, levelQ, levelP)

    def NewPoly(self, batch=None) -> Poly:
        return Poly(self.RingQ.NewPoly(batch), self.RingP.NewPoly(batch) if self.RingP is not None else None)

    def _both(self, name, *polys):
        getattr(self.RingQ, name)(*[p.Q for p in polys])
        if self.RingP is not None:
            getattr(self.RingP, name)(*[p.P for p in polys])

    def AutomorphismNTTWithIndex(self, p1: Poly, index, p2: Poly):            # ring/ringqp/operations.go:302
        self.RingQ.AutomorphismNTTWithIndex(p1.Q, index, p2.Q)
        if self.RingP is not None: self.RingP.AutomorphismNTTWithIndex(p1.P, index, p2.P)

    def AutomorphismNTTWithIndexThenAddLazy(self, p1: Poly, index, p2: Poly):  # ring/ringqp/operations.go:314
        self.RingQ.AutomorphismNTTWithIndexThenAddLazy(p1.Q, index, p2.Q)
        if self.RingP is not None: self.RingP.AutomorphismNTTWithIndexThenAddLazy(p1.P, index, p2.P)

    def AutomorphismNTT(self, p1: Poly, galEl: int, p2: Poly):                # ring/ringqp/operations.go:290
        self.RingQ.AutomorphismNTT(p1.Q, galEl, p2.Q)
        if self.RingP is not None: self.RingP.AutomorphismNTT(p1.P, galEl, p2.P)

    def Automorphism(self, p1: Poly, galEl: int, p2: Poly):                   # ring/ringqp/operations.go:278
        self.RingQ.Automorphism(p1.Q, galEl, p2.Q)
        if self.RingP is not None: self.RingP.Automorphism(p1.P, galEl, p2.P)

    def MulScalar(self, p1: Poly, scalar: int, p2: Poly):                     # ring/ringqp/operations.go:105
        self.RingQ.MulScalar(p1.Q, scalar, p2.Q)
        if self.RingP is not None: self.RingP.MulScalar(p1.P, scalar, p2.P)

    def MulRNSScalarMontgomery(self, p: Poly, scalar, pOut: Poly):            # ring/ringqp/operations.go:244
        nq = len(self.ctx.Q)
        self.RingQ.MulRNSScalarMontgomery(p.Q, list(scalar[:nq]), pOut.Q)
        if self.RingP is not None: self.RingP.MulRNSScalarMontgomery(p.P, list(scalar[nq:]), pOut.P)

    def EvalPolyScalar(self, pol, pt: int, p3: Poly):                         # ring/ringqp/operations.go:92
        self.RingQ.EvalPolyScalar([p.Q for p in pol], pt, p3.Q)
        if self.RingP is not None: self.RingP.EvalPolyScalar([p.P for p in pol], pt, p3.P)

    def ExtendBasisSmallNormAndCenter(self, polyInQ, levelP: int, polyOutQ, polyOutP):   # ring/ringqp/operations.go:325
        import ctypes
        from . import _lib
        from .ring import _dptr, _stream
        if polyOutQ.data_ptr() != polyInQ.data_ptr():
            polyOutQ.copy_(polyInQ)
        b = polyInQ.shape[0] if polyInQ.dim() == 3 else 1
        sq = polyInQ.shape[-2] * polyInQ.shape[-1] if polyInQ.dim() == 3 else 0
        sp = polyOutP.shape[-2] * polyOutP.shape[-1] if polyOutP.dim() == 3 else 0
        _lib.check(_lib.lib().lgpu_extend_basis_small_norm_and_center(self.ctx.h, _dptr(polyInQ), levelP, _dptr(polyOutP), b, sq, sp, _stream()))

    def __getattr__(self, name):
        # Add, Sub, Neg, Reduce, MForm, IMForm, MulCoeffsMontgomery[Lazy|ThenAdd|...], NTT, INTT, ...
        if name in OPS or name in ("NTT", "NTTLazy", "INTT", "INTTLazy"):
            return lambda *polys: self._both(name, *polys)
        raise AttributeError(name)
