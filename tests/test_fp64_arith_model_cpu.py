"""CPU model of the FP64-pipe modular arithmetic of the device kernels (lattigo_b200/csrc/ntt_arith.cuh: fp_mulmod, fp_reduce, fp_canon;
keyswitch_fused.cu: ks_ext_split, the FP64-pipe MAC of ks_chunk_mac_fp8r_kernel<1>), evaluated with correctly rounded IEEE operations
(fused multiply-add emulated through exact rationals) on the WORST-CASE operands the bounds in the kernel comments allow, and compared with
exact integer arithmetic. The GPU parity tests exercise the same code on random data; this pins the range analysis at its edges:
  * q up to the fp_ok limit ((10 + logN) q < 2^51), lazy butterfly operands up to (10 + logN) q,
  * basis-extension sums over 6 sources of maximal residues and constants,
  * 32 accumulated MAC terms."""
import random
from fractions import Fraction

import pytest

MAGIC = 6755399441055744.0          # 1.5 * 2^52
TWO52 = 4503599627370496.0


def fma(a, b, c):
    return float(Fraction(a) * Fraction(b) + Fraction(c))      # one rounding, like the hardware instruction


def fp_mulmod(v, w, q, qinv):
    h = v * w
    lo = fma(v, w, -h)
    t = fma(h, qinv, MAGIC) - MAGIC
    r = fma(-t, q, h)
    return r + lo


def fp_reduce(x, q, qinv):
    t = fma(x, qinv, MAGIC) - MAGIC
    return fma(-t, q, x)


def fp_canon(x, q, qinv):
    r = fp_reduce(x, q, qinv)
    return int(r + q) if r < 0.0 else int(r)


def primes_near_limit(logN):
    """a few odd moduli just below the fp_ok limit and a small one (primality is irrelevant to the range analysis)"""
    lim = ((1 << 51) - 1) // (10 + logN)
    top = lim if lim & 1 else lim - 1
    return [top - 2 * i for i in range(3)] + [(1 << 45) + 59, (1 << 36) + 31, (1 << 20) + 7]


@pytest.mark.parametrize("logN", [13, 16])
def test_fp_mulmod_and_canon_on_the_lazy_range(logN):
    rng = random.Random(1)
    for q in primes_near_limit(logN):
        fq, fqinv = float(q), 1.0 / float(q)
        bound = (10 + logN) * q - 1
        for v in [bound, -bound, q - 1, 1 - q, 0] + [rng.randrange(-bound, bound + 1) for _ in range(200)]:
            for w in (q - 1, 1, rng.randrange(q)):
                r = fp_mulmod(float(v), float(w), fq, fqinv)
                assert r == int(r) and abs(r) <= q, (q, v, w, r)                     # an exact integer, within one modulus of zero
                assert (int(r) - v * w) % q == 0
                assert fp_canon(r, fq, fqinv) == (v * w) % q
        for x in (bound, -bound, 0, rng.randrange(-bound, bound)):
            assert fp_canon(float(x), fq, fqinv) == x % q


def ks_ext_split(ys, cs, vt, half, q):
    """device sequence of keyswitch_fused.cu:ks_ext_split for one coefficient (sources ys, plain constants cs)"""
    fq, fqinv = float(q), 1.0 / float(q)
    w = float(fp_canon(8388608.0, fq, fqinv))
    A = B = 0
    for y, c in zip(ys, cs):
        d = fp_canon(fp_mulmod(float(c), w, fq, fqinv), fq, fqinv)       # c * 2^23 mod q, computed when the tile is staged
        assert d == (c << 23) % q
        y0, y1, c0, c1, d0, d1 = y & 0x7FFFFF, y >> 23, c & 0x7FFFFF, c >> 23, d & 0x7FFFFF, d >> 23
        assert max(y1, c1, d1) < (1 << 32)
        A += y0 * c0 + y1 * d0
        B += y0 * c1 + y1 * d1
    assert A < (1 << 52) and B < (1 << 52)                                 # the biased conversions are exact
    ab, bb = TWO52 + float(A), TWO52 + float(B)
    vtb = (float(vt) - float(half)) - TWO52
    t = fp_reduce(fma(bb, 8388608.0, -37778931862957161709568.0), fq, fqinv)
    assert fma(bb, 8388608.0, -37778931862957161709568.0) == float(B << 23)
    return fp_reduce((ab + vtb) + t, fq, fqinv)


@pytest.mark.parametrize("logN", [13, 16])
def test_ks_ext_split_equals_the_exact_basis_extension(logN):
    rng = random.Random(2)
    for q in primes_near_limit(logN):
        srcq = primes_near_limit(logN)[:3] * 2                               # six source moduli at the limit
        cases = [([s - 1 for s in srcq], [q - 1] * 6)] + [([rng.randrange(s) for s in srcq], [rng.randrange(q) for _ in srcq]) for _ in range(100)]
        for ys, cs in cases:
            for vt, half in ((q - 1, 0), (0, (q - 1) // 2), (rng.randrange(q), q // 2)):
                r = ks_ext_split(ys, cs, vt, half, q)
                assert r == int(r) and abs(r) <= 0.66 * q, (q, r)
                assert (int(r) - (sum(y * c for y, c in zip(ys, cs)) + vt - half)) % q == 0


@pytest.mark.parametrize("logN", [13, 16])
def test_fp64_mac_of_the_key_switch(logN):
    """acc = sum_d fp_mulmod(x_d, key_d) over up to 32 digits, then one fp_mulmod by 2^-64 mod q: the canonical residue of
    sum_d x_d key_d 2^-64, i.e. what sum_d MRedLazy(key_d, x_d) reduces to (core/rlwe/evaluator_gadget_product.go:129-201)."""
    rng = random.Random(3)
    for q in primes_near_limit(logN):
        fq, fqinv = float(q), 1.0 / float(q)
        rinv = pow(1 << 64, -1, q)
        bound = (10 + logN) * q - 1
        for nd in (1, 11, 32):
            for trial in range(20):
                xs = [bound if trial == 0 else -bound if trial == 1 else rng.randrange(-bound, bound + 1) for _ in range(nd)]
                ks = [(1 << 46) - 1 if trial < 2 else rng.randrange(q) for _ in range(nd)]       # the guard admits key words below 2^46
                acc = 0.0
                for x, k in zip(xs, ks):
                    acc = acc + fp_mulmod(float(x), float(k), fq, fqinv)
                    assert acc == int(acc) and abs(acc) < TWO52
                got = fp_canon(fp_mulmod(acc, float(rinv), fq, fqinv), fq, fqinv)
                assert got == (sum(x * k for x, k in zip(xs, ks)) * rinv) % q
