"""Scalar restatement of the CKKS encoder's special FFT -- `schemes/ckks/ckks_vector_ops.go:18-77` (SpecialIFFTDouble, SpecialFFTDouble) and the
tables of `schemes/ckks/encoder.go:83-112` / `schemes/ckks/utils.go` (rotGroup = powers of 5 modulo m, roots = e^{2 pi i k / m}).
TEST INFRASTRUCTURE ONLY. Python floats are IEEE doubles without fusion, and the complex products are written out the way Go evaluates them
((ac - bd) + (ad + bc) i), so this is bit-compatible with the reference on amd64; pinned by the transform identities in
tests/test_oracle_specialfft.py (IFFT o FFT = id to rounding, and FFT = the decoding matrix on the rotation group)."""
from __future__ import annotations

import math


def rot_group(m: int):
    out, f = [], 1
    for _ in range(m >> 2):
        out.append(f)
        f = f * 5 & (m - 1)
    return out


def roots(m: int):
    """e^{2 pi i k / m}, k = 0..m (what GetRootsComplex128, schemes/ckks/utils.go, tabulates). The device is handed whatever table the caller
    uses, so parity does not depend on how the table itself is rounded."""
    angle = 2 * math.pi / m
    return [complex(math.cos(angle * i), math.sin(angle * i)) for i in range(m + 1)]


def _mul(a, b):
    return complex(a.real * b.real - a.imag * b.imag, a.real * b.imag + a.imag * b.real)


def _bitrev(v, n):
    log = n.bit_length() - 1
    for i in range(n):
        j = int(format(i, "0%db" % log)[::-1], 2) if log else 0
        if j > i:
            v[i], v[j] = v[j], v[i]


def special_fft(values, n, m, rg, rt):
    _bitrev(values, n)
    logn, logm = (n - 1).bit_length(), (m - 1).bit_length()
    for loglen in range(1, logn + 1):
        ln = 1 << loglen; lenh = ln >> 1; lenq = ln << 2
        gap = logm - 2 - loglen; mask = lenq - 1
        for i in range(0, n, ln):
            for j in range(lenh):
                k = i + j
                values[k + lenh] = _mul(values[k + lenh], rt[(rg[j] & mask) << gap])
                a, b = values[k], values[k + lenh]
                values[k], values[k + lenh] = complex(a.real + b.real, a.imag + b.imag), complex(a.real - b.real, a.imag - b.imag)


def special_ifft(values, n, m, rg, rt):
    logn, logm = (n - 1).bit_length(), (m - 1).bit_length()
    for loglen in range(logn, 0, -1):
        ln = 1 << loglen; lenh = ln >> 1; lenq = ln << 2
        gap = logm - 2 - loglen; mask = lenq - 1
        for i in range(0, n, ln):
            for j in range(lenh):
                k = i + j
                a, b = values[k], values[k + lenh]
                d = complex(a.real - b.real, a.imag - b.imag)
                values[k] = complex(a.real + b.real, a.imag + b.imag)
                values[k + lenh] = _mul(d, rt[(lenq - (rg[j] & mask)) << gap])
    for i in range(n):
        values[i] = complex(values[i].real / n, values[i].imag / n)
    _bitrev(values, n)
