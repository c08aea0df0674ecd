"""Host-side mirror of the reference's `ring.Ring` / `ring.SubRing` method set, bound to the C ABI.

Polynomials are CUDA torch tensors of dtype int64 (bit pattern of uint64), shape (limbs, N) or
(batch, limbs, N), contiguous: the (limb, coeff) row-major layout of include/lattigo_b200.h. Method names,
argument order and in-place rules follow the reference (ring/operations.go, ring/ntt.go,
ring/subring_ops.go); every call is asynchronous on torch's current CUDA stream."""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import numpy as np

from . import _lib

RING_Q, RING_P = 0, 1

# Opcode numbering of include/lattigo_b200.h (enum lgpu_vecop_code), named after the SubRing methods.
OPS = [
    "Add", "AddLazy", "Sub", "SubLazy", "Neg", "Reduce", "ReduceLazy",
    "MulCoeffsLazy", "MulCoeffsLazyThenAddLazy",
    "MulCoeffsBarrett", "MulCoeffsBarrettLazy", "MulCoeffsBarrettThenAdd", "MulCoeffsBarrettThenAddLazy",
    "MulCoeffsMontgomery", "MulCoeffsMontgomeryLazy", "MulCoeffsMontgomeryThenAdd",
    "MulCoeffsMontgomeryThenAddLazy", "MulCoeffsMontgomeryLazyThenAddLazy",
    "MulCoeffsMontgomeryThenSub", "MulCoeffsMontgomeryThenSubLazy", "MulCoeffsMontgomeryLazyThenSubLazy",
    "MulCoeffsMontgomeryLazyThenNeg",
    "AddLazyThenMulScalarMontgomery", "AddScalarLazyThenMulScalarMontgomery",
    "AddScalar", "AddScalarLazy", "AddScalarLazyThenNegTwoModulusLazy", "SubScalar",
    "MulScalarMontgomery", "MulScalarMontgomeryLazy", "MulScalarMontgomeryThenAdd",
    "MulScalarMontgomeryThenAddScalar", "SubThenMulScalarMontgomeryTwoModulus",
    "MForm", "MFormLazy", "IMForm", "Zero", "Mask",
]
OP = {n: k for k, n in enumerate(OPS)}


def _torch():
    import torch
    return torch


def _stream():
    return ctypes.c_void_p(_torch().cuda.current_stream().cuda_stream)


def _dptr(t):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous() and t.element_size() == 8, "polynomials must be contiguous 64-bit CUDA tensors"
    return ctypes.c_void_p(t.data_ptr())


def _u64arr(vals: Optional[Sequence[int]]):
    if vals is None:
        return None, None
    a = np.array([int(v) & 0xFFFFFFFFFFFFFFFF for v in vals], dtype=np.uint64)
    return a, ctypes.c_void_p(a.ctypes.data)


class Context:
    """Owns an lgpu_ctx: ringQ, ringP, BasisExtender and Decomposer constants for one (N, Q, P)
    (ring.NewRing + ring.NewBasisExtender + ring.NewDecomposer). device < 0 builds a host-only
    context (tables, no device) for CPU-side checks."""

    def __init__(self, logN: int, Q: Sequence[int], P: Sequence[int] = (), device: int = 0, ring_type: int = 0):
        self.logN, self.N = logN, 1 << logN
        self.Q = [int(x) for x in Q]
        self.P = [int(x) for x in P]
        self.device = device
        self.ring_type = ring_type        # 0 = ring.Standard, 1 = ring.ConjugateInvariant (NthRoot = 4N)
        qa = np.array(self.Q, dtype=np.uint64)
        pa = np.array(self.P, dtype=np.uint64) if self.P else None
        h = ctypes.c_void_p()
        _lib.check(_lib.lib().lgpu_create(ctypes.byref(h), device, logN, ring_type, qa.ctypes.data, len(self.Q),
                                          pa.ctypes.data if pa is not None else None, len(self.P)))
        self.h = h
        self.ringQ = Ring(self, RING_Q, len(self.Q) - 1)
        self.ringP = Ring(self, RING_P, len(self.P) - 1) if self.P else None

    def close(self):
        if getattr(self, "h", None):
            _lib.lib().lgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        _lib.check(_lib.lib().lgpu_sync(self.h, _stream() if self.device >= 0 else None))

    # table read-back (host)
    def table(self, ring: int, limb: int, kind: int) -> np.ndarray:
        half = self.N << (1 if self.ring_type else 0)            # NthRoot / 2 root-table entries
        n = {0: 6, 1: half, 2: half, 3: max(limb, 1)}[kind]
        out = np.zeros(n, dtype=np.uint64)
        _lib.check(_lib.lib().lgpu_ring_get_table(self.h, ring, limb, kind, out.ctypes.data, n))
        return out

    # tensors
    def new_poly(self, limbs: int, batch: Optional[int] = None):
        torch = _torch()
        shape = (limbs, self.N) if batch is None else (batch, limbs, self.N)
        return torch.zeros(shape, dtype=torch.int64, device="cuda:%d" % self.device)

    def to_device(self, a: np.ndarray):
        torch = _torch()
        return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).to("cuda:%d" % self.device)

    @staticmethod
    def to_host(t) -> np.ndarray:
        return t.detach().cpu().numpy().view(np.uint64)


class SubRing:
    """ring.SubRing (ring/subring.go:15-35): single-row operations (ring/subring_ops.go:6-273)."""

    def __init__(self, ctx: Context, which: int, limb: int):
        self.ctx, self.which, self.limb = ctx, which, limb
        self.Modulus = (ctx.Q if which == RING_Q else ctx.P)[limb]
        self.N = ctx.N

    def _vec(self, name, p1, p2, p3, s0=0, s1=0):
        n = p3.shape[-1]
        _lib.check(_lib.lib().lgpu_subring_vecop(self.ctx.h, self.which, self.limb, OP[name], _dptr(p1), _dptr(p2), _dptr(p3),
                                                 int(s0) & 0xFFFFFFFFFFFFFFFF, int(s1) & 0xFFFFFFFFFFFFFFFF, n, _stream()))

    def NTT(self, p1, p2): _lib.check(_lib.lib().lgpu_subring_ntt(self.ctx.h, self.which, self.limb, _dptr(p1), _dptr(p2), 0, _stream()))
    def NTTLazy(self, p1, p2): _lib.check(_lib.lib().lgpu_subring_ntt(self.ctx.h, self.which, self.limb, _dptr(p1), _dptr(p2), 1, _stream()))
    def INTT(self, p1, p2): _lib.check(_lib.lib().lgpu_subring_intt(self.ctx.h, self.which, self.limb, _dptr(p1), _dptr(p2), 0, _stream()))
    def INTTLazy(self, p1, p2): _lib.check(_lib.lib().lgpu_subring_intt(self.ctx.h, self.which, self.limb, _dptr(p1), _dptr(p2), 1, _stream()))

    def __getattr__(self, name):
        if name in OP:
            two_in = {"Add", "AddLazy", "Sub", "SubLazy"} | {n for n in OPS if n.startswith("MulCoeffs")}
            if name in two_in:
                return lambda p1, p2, p3: self._vec(name, p1, p2, p3)
            if name in ("Neg", "Reduce", "ReduceLazy", "MForm", "MFormLazy", "IMForm"):
                return lambda p1, p2: self._vec(name, p1, None, p2)
            if name == "AddLazyThenMulScalarMontgomery":
                return lambda p1, p2, scalarMont, p3: self._vec(name, p1, p2, p3, scalarMont)
            if name == "SubThenMulScalarMontgomeryTwoModulus":
                return lambda p1, p2, scalarMont, p3: self._vec(name, p1, p2, p3, scalarMont)
            if name in ("AddScalarLazyThenMulScalarMontgomery",):
                return lambda p1, scalar0, scalarMont1, p2: self._vec(name, p1, None, p2, scalar0, scalarMont1)
            if name == "MulScalarMontgomeryThenAddScalar":
                return lambda p1, scalar0, scalarMont1, p2: self._vec(name, p1, None, p2, scalar0, scalarMont1)
            if name == "Mask":
                return lambda p1, w, mask, p2: self._vec(name, p1, None, p2, w, mask)
            if name == "Zero":
                return lambda p1: self._vec(name, None, None, p1)
            # single-scalar forms: (p1, scalar, p2)
            return lambda p1, scalar, p2: self._vec(name, p1, None, p2, scalar)
        raise AttributeError(name)


class Ring:
    """ring.Ring at a level (ring/ring.go:71-82, AtLevel :184-209)."""

    def __init__(self, ctx: Context, which: int, level: int):
        self.ctx, self.which, self.level = ctx, which, level
        self.moduli = ctx.Q if which == RING_Q else ctx.P

    def AtLevel(self, level: int) -> "Ring":
        if level < 0:
            raise _lib.LgpuError("level cannot be negative")
        if level > len(self.moduli) - 1:
            raise _lib.LgpuError("level cannot be larger than max level")
        return Ring(self.ctx, self.which, level)

    def N(self): return self.ctx.N
    def Level(self): return self.level
    def MaxLevel(self): return len(self.moduli) - 1
    def ModuliChain(self): return list(self.moduli)
    @property
    def SubRings(self): return [SubRing(self.ctx, self.which, i) for i in range(len(self.moduli))]
    def NewPoly(self, batch=None): return self.ctx.new_poly(self.level + 1, batch)

    def _batch(self, t):
        if t.dim() == 3:
            assert t.shape[1] >= self.level + 1
            return t.shape[0], t.shape[1] * t.shape[2]
        assert t.shape[0] >= self.level + 1
        return 1, 0

    def _same_layout(self, ref, *others):
        """The C ABI applies ONE (batch, batch stride) to every operand of a call: batched operands of another limb count
        (polynomials of different levels under AtLevel) would be read with the wrong stride."""
        for t in others:
            if t is not None and t.dim() == 3 and ref.dim() == 3:
                assert t.shape[0] == ref.shape[0] and t.shape[1] == ref.shape[1], \
                    "batched operands must share the batch size and the limb count of the output (one batch stride per call)"
            elif t is not None:
                assert t.dim() == ref.dim(), "operands must be all batched or all single polynomials"

    def _ntt(self, inverse, p1, p2, lazy):
        b, bs = self._batch(p2)
        self._same_layout(p2, p1)
        f = _lib.lib().lgpu_intt if inverse else _lib.lib().lgpu_ntt
        _lib.check(f(self.ctx.h, self.which, self.level, _dptr(p1), _dptr(p2), lazy, b, bs, _stream()))

    def NTTThenMulCoeffsMontgomery(self, p1, p2, p3):
        """p3 = MulCoeffsMontgomery(NTT(p1), p2): ring/ntt.go:127-131 + ring/operations.go:88-92 in one pass over HBM."""
        b, bs = self._batch(p3)
        self._same_layout(p3, p1, p2)
        _lib.check(_lib.lib().lgpu_ntt_then_mul_coeffs_montgomery(self.ctx.h, self.which, self.level, _dptr(p1), _dptr(p2), _dptr(p3), b, bs, _stream()))

    def NTT(self, p1, p2): self._ntt(False, p1, p2, 0)          # ring/ntt.go:127
    def NTTLazy(self, p1, p2): self._ntt(False, p1, p2, 1)      # ring/ntt.go:134
    def INTT(self, p1, p2): self._ntt(True, p1, p2, 0)          # ring/ntt.go:141
    def INTTLazy(self, p1, p2): self._ntt(True, p1, p2, 1)      # ring/ntt.go:148

    def _vec(self, name, p1, p2, p3, s0=None, s1=None):
        b, bs = self._batch(p3)
        self._same_layout(p3, p1, p2)
        a0, p0 = _u64arr(s0)
        a1, pp1 = _u64arr(s1)
        _lib.check(_lib.lib().lgpu_vecop(self.ctx.h, self.which, self.level, OP[name], _dptr(p1), _dptr(p2), _dptr(p3),
                                         p0, pp1, b, bs, _stream()))

    def __getattr__(self, name):
        if name in OP:
            two_in = {"Add", "AddLazy", "Sub", "SubLazy"} | {n for n in OPS if n.startswith("MulCoeffs")}
            if name in two_in:
                return lambda p1, p2, p3: self._vec(name, p1, p2, p3)
            if name in ("Neg", "Reduce", "ReduceLazy", "MForm", "MFormLazy", "IMForm"):
                return lambda p1, p2: self._vec(name, p1, None, p2)
        raise AttributeError(name)

    def _lvl_mods(self): return self.moduli[: self.level + 1]

    def AddScalar(self, p1, scalar: int, p2):       # ring/operations.go:151
        self._vec("AddScalar", p1, None, p2, [scalar % q for q in self._lvl_mods()])

    def SubScalar(self, p1, scalar: int, p2):       # ring/operations.go:186
        self._vec("SubScalar", p1, None, p2, [scalar % q for q in self._lvl_mods()])

    AddScalarBigint = AddScalar                     # ring/operations.go:158 (python ints are big ints)
    SubScalarBigint = SubScalar                     # ring/operations.go:193

    def MulScalar(self, p1, scalar: int, p2):       # ring/operations.go:201
        self._vec("MulScalarMontgomery", p1, None, p2, [((scalar % q) << 64) % q for q in self._lvl_mods()])

    def MulScalarThenAdd(self, p1, scalar: int, p2):   # ring/operations.go:208
        self._vec("MulScalarMontgomeryThenAdd", p1, None, p2, [((scalar % q) << 64) % q for q in self._lvl_mods()])

    MulScalarBigint = MulScalar                     # ring/operations.go:231
    MulScalarBigintThenAdd = MulScalarThenAdd       # ring/operations.go:240

    def MulScalarThenSub(self, p1, scalar: int, p2):   # ring/operations.go:223
        self._vec("MulScalarMontgomeryThenAdd", p1, None, p2, [((q - scalar % q) << 64) % q for q in self._lvl_mods()])

    def MulRNSScalarMontgomery(self, p1, scalar: Sequence[int], p2):   # ring/operations.go:216
        self._vec("MulScalarMontgomery", p1, None, p2, list(scalar)[: self.level + 1])

    def _halves(self, name, p1, s0, s1, p2):
        # the *DoubleRNSScalar family: one SubRing op on coefficients [0, N/2), another on [N/2, N)
        assert p2.dim() == 2, "DoubleRNSScalar ops take a single (limbs, N) polynomial"
        h = self.ctx.N >> 1
        for i, sr in enumerate(self.SubRings[: self.level + 1]):
            sr._vec(name, p1[i, :h], None, p2[i, :h], s0[i])
            sr._vec(name, p1[i, h:], None, p2[i, h:], s1[i])

    def AddDoubleRNSScalar(self, p1, scalar0, scalar1, p2):        # ring/operations.go:167
        self._halves("AddScalar", p1, scalar0, scalar1, p2)

    def SubDoubleRNSScalar(self, p1, scalar0, scalar1, p2):        # ring/operations.go:177
        self._halves("SubScalar", p1, scalar0, scalar1, p2)

    def MulDoubleRNSScalar(self, p1, scalar0, scalar1, p2):        # ring/operations.go:250
        m = self._lvl_mods()
        self._halves("MulScalarMontgomery", p1, [(int(a) << 64) % q for a, q in zip(scalar0, m)], [(int(a) << 64) % q for a, q in zip(scalar1, m)], p2)

    def MulDoubleRNSScalarThenAdd(self, p1, scalar0, scalar1, p2):  # ring/operations.go:260
        m = self._lvl_mods()
        self._halves("MulScalarMontgomeryThenAdd", p1, [(int(a) << 64) % q for a, q in zip(scalar0, m)], [(int(a) << 64) % q for a, q in zip(scalar1, m)], p2)

    def EvalPolyScalar(self, p1: Sequence, scalar: int, p2):       # ring/operations.go:269
        p2[..., : self.level + 1, :].copy_(p1[-1][..., : self.level + 1, :])
        for i in range(len(p1) - 1, 0, -1):
            self.MulScalar(p2, scalar, p2)
            self.Add(p2, p1[i - 1], p2)

    def Shift(self, p1, k: int, p2):                # ring/operations.go:278
        b, bs = self._batch(p2)
        _lib.check(_lib.lib().lgpu_shift(self.ctx.h, self.which, self.level, _dptr(p1), int(k), _dptr(p2), b, bs, _stream()))

    def MultByMonomial(self, p1, k: int, p2):       # ring/operations.go:306
        b, bs = self._batch(p2)
        _lib.check(_lib.lib().lgpu_mult_by_monomial(self.ctx.h, self.which, self.level, _dptr(p1), int(k), _dptr(p2), b, bs, _stream()))

    def MulByVectorMontgomery(self, p1, vector, p2):               # ring/operations.go:366
        for i, sr in enumerate(self.SubRings[: self.level + 1]): sr._vec("MulCoeffsMontgomery", p1[i], vector, p2[i])

    def MulByVectorMontgomeryThenAddLazy(self, p1, vector, p2):    # ring/operations.go:373
        for i, sr in enumerate(self.SubRings[: self.level + 1]): sr._vec("MulCoeffsMontgomeryThenAddLazy", p1[i], vector, p2[i])

    # ring/automorphism.go and ring/scaling.go (thin wrappers over the C ABI; batch = leading dimension)
    def AutomorphismNTTIndex(self, galEl: int):     # ring/automorphism.go:12
        idx = self.ctx.new_poly(1)[0]
        _lib.check(_lib.lib().lgpu_automorphism_ntt_index(self.ctx.h, int(galEl), _dptr(idx), _stream()))
        return idx

    def AutomorphismNTTWithIndex(self, polIn, index, polOut):      # ring/automorphism.go:50
        b, bs = self._batch(polOut)
        _lib.check(_lib.lib().lgpu_automorphism_ntt_with_index(self.ctx.h, self.which, self.level, _dptr(polIn), _dptr(index), _dptr(polOut), 0, b, bs, _stream()))

    def AutomorphismNTTWithIndexThenAddLazy(self, polIn, index, polOut):   # ring/automorphism.go:82
        b, bs = self._batch(polOut)
        _lib.check(_lib.lib().lgpu_automorphism_ntt_with_index(self.ctx.h, self.which, self.level, _dptr(polIn), _dptr(index), _dptr(polOut), 1, b, bs, _stream()))

    def AutomorphismNTT(self, polIn, galEl: int, polOut):          # ring/automorphism.go:38
        b, bs = self._batch(polOut)
        _lib.check(_lib.lib().lgpu_automorphism_ntt(self.ctx.h, self.which, self.level, _dptr(polIn), int(galEl), _dptr(polOut), b, bs, _stream()))

    def Automorphism(self, polIn, galEl: int, polOut):             # ring/automorphism.go:113
        b, bs = self._batch(polOut)
        _lib.check(_lib.lib().lgpu_automorphism(self.ctx.h, self.which, self.level, _dptr(polIn), int(galEl), _dptr(polOut), b, bs, _stream()))


def MapSmallDimensionToLargerDimensionNTT(ctx: "Context", polSmall, polLarge):   # ring/operations.go:380
    rows = min(polSmall.shape[0], polLarge.shape[0])
    assert polSmall.dim() == 2 and polLarge.dim() == 2
    _lib.check(_lib.lib().lgpu_map_small_dimension_to_larger_dimension_ntt(ctx.h, _dptr(polSmall), polSmall.shape[1], _dptr(polLarge), polLarge.shape[1],
                                                                          rows, _stream()))
