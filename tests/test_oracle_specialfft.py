"""oracle/specialfft.py against what the transform is: FFT(x)[i] = sum_j x[j] * zeta^(5^i * j) over the rotation group (the CKKS decoding map,
schemes/ckks/encoder.go:83-112), and IFFT is its inverse."""
import cmath
import math

from oracle import specialfft as SF


def test_special_fft_is_the_decoding_matrix_and_ifft_inverts_it():
    for logn in (1, 3, 5):
        n = 1 << logn
        m = 4 * n * 2                       # a ring with more slots than used: m = 2N, n <= N/2
        rg, rt = SF.rot_group(m), SF.roots(m)
        x = [complex(math.sin(3 * i + 1), math.cos(5 * i)) for i in range(n)]
        y = list(x)
        SF.special_fft(y, n, m, rg, rt)
        # evaluation of the polynomial sum_j x[j] X^j at the points zeta_{4n}^{5^i}
        for i in range(n):
            z = cmath.exp(2j * math.pi * (rg[i] % (4 * n)) / (4 * n))
            ref = sum(x[j] * z ** j for j in range(n))
            assert abs(ref - y[i]) < 1e-9 * n, (logn, i)
        SF.special_ifft(y, n, m, rg, rt)
        assert max(abs(a - b) for a, b in zip(x, y)) < 1e-12
