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

    def __getattr__(self, name):
        # Add, Sub, Neg, Reduce, MForm, IMForm, MulCoeffsMontgomery[Lazy|ThenAdd|...], NTT, INTT, ...
        if name in OPS or name in ("NTT", "NTTLazy", "INTT", "INTTLazy"):
            return lambda *polys: self._both(name, *polys)
        raise AttributeError(name)
