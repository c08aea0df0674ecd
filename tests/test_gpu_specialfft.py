"""lgpu_ckks_special_fft against oracle/specialfft.py (schemes/ckks/ckks_vector_ops.go:18-77) with tolerance ZERO: the device evaluates the
reference's floating-point expressions in the reference's order without fused multiply-add, on the same root / rotation tables, so every
double must come out bit-identical (compared through their 64-bit patterns). Sizes: inside one shared-memory chunk, exactly one chunk,
and several chunks plus wide stages (2^13 slots); batch of 3."""
import numpy as np
import pytest

from oracle import specialfft as SF

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("logn,logN_ring", [(3, 8), (11, 12), (13, 14), (5, 16)])
def test_special_fft_and_ifft_bit_exact(logn, logN_ring):
    import torch
    import lattigo_b200 as lb
    from oracle import oracle as O
    n, m = 1 << logn, 2 << logN_ring
    rg, rt = SF.rot_group(m), SF.roots(m)
    rng = np.random.default_rng(logn)
    batch = 3
    x = rng.normal(size=(batch, n)) + 1j * rng.normal(size=(batch, n))
    q, _ = O.gen_moduli(9, [40], [])
    ctx = lb.Context(8, q)                                     # the transform does not depend on the ring; any context provides the device
    try:
        d_rg = torch.tensor(rg, dtype=torch.int64, device="cuda")
        d_rt = torch.tensor(np.array(rt, dtype=np.complex128), device="cuda")
        for inverse in (False, True):
            want = []
            for b in range(batch):
                v = [complex(z) for z in x[b]]
                (SF.special_ifft if inverse else SF.special_fft)(v, n, m, rg, rt)
                want.append(np.array(v, dtype=np.complex128))
            want = np.stack(want)
            d = torch.tensor(x, dtype=torch.complex128, device="cuda")
            lb.ckks_fft.special_fft(ctx, d, m, d_rg, d_rt, inverse)
            got = d.cpu().numpy()
            assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), (logn, inverse, float(np.abs(got - want).max()))
        with pytest.raises(lb.LgpuError, match="powers of two"):
            lb.ckks_fft.special_fft(ctx, torch.zeros((1, 24), dtype=torch.complex128, device="cuda"), m, d_rg, d_rt)
    finally:
        ctx.close()
