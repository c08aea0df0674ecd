"""ckks.SpecialFFTDouble / SpecialIFFTDouble (schemes/ckks/ckks_vector_ops.go:18-77) bound to the C ABI: in place on (batch, n) complex128
CUDA tensors with the encoder's own rotGroup / roots tables (device tensors)."""
from __future__ import annotations

import ctypes

from . import _lib
from .ring import Context, _stream


def special_fft(ctx: Context, values, m: int, rot_group, roots, inverse: bool = False):
    import torch
    assert values.dtype == torch.complex128 and values.is_cuda and values.is_contiguous()
    assert rot_group.dtype == torch.int64 and roots.dtype == torch.complex128
    batch, n = (values.shape[0], values.shape[1]) if values.dim() == 2 else (1, values.shape[0])
    _lib.check(_lib.lib().lgpu_ckks_special_fft(ctx.h, ctypes.c_void_p(values.data_ptr()), n, m, ctypes.c_void_p(rot_group.data_ptr()),
                                                ctypes.c_void_p(roots.data_ptr()), 1 if inverse else 0, batch, n, _stream()))
    return values
