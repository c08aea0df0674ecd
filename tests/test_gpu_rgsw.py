"""GPU parity of lgpu_rgsw_external_product against oracle/rgsw.py (core/rgsw/evaluator.go:39-283), bit for bit, on the three
reference code paths and on the fused key-switch sizes; uniform operands (parity is over Z_q identities)."""
import numpy as np
import pytest

from oracle import oracle as O
from oracle import rgsw as RG
from tests import helpers as H

pytestmark = pytest.mark.gpu
U64 = np.uint64


def _case(logN, q, p, pw2, seed, batch=2, ct_extra_levels=0):
    import torch
    import lattigo_b200 as lb
    params_key = O.Parameters(logN, q[: len(q) - ct_extra_levels], p)         # the RGSW ciphertext may sit below the RLWE ciphertext
    N = params_key.N()
    rng = np.random.default_rng(seed)
    levelQ, levelP = params_key.MaxLevelQ(), params_key.MaxLevelP()
    rg_o = [H.random_gadget_ciphertext(params_key, levelQ, levelP, rng, pw2=pw2) for _ in range(2)]
    ctx = lb.Context(logN, q, p)
    try:
        rg_d = lb.rgsw.Ciphertext(*[lb.GadgetCiphertext(ctx, g.data, levelQ, levelP, pw2, g.pw2_sizes) for g in rg_o])
        ct = np.stack([np.stack([H.rand_poly(q, N, rng), H.rand_poly(q, N, rng)]) for _ in range(batch)])
        ev_o = RG.Evaluator(params_key)
        want = np.zeros((batch, 2, levelQ + 1, N), dtype=U64)
        for b in range(batch):
            ev_o.ExternalProduct([ct[b, 0, : levelQ + 1].copy(), ct[b, 1, : levelQ + 1].copy()], rg_o, [want[b, 0], want[b, 1]])
        ev = lb.rgsw.Evaluator(ctx)
        d = ctx.to_device(ct)
        out = torch.zeros_like(d)
        ev.ExternalProduct(d, rg_d, out)
        assert np.array_equal(ctx.to_host(out)[:, :, : levelQ + 1], want), (logN, pw2)
        if ct_extra_levels:
            assert int(out[:, :, levelQ + 1:].abs().max()) == 0
        ev.ExternalProduct(d, rg_d, d)                                        # in place (rgsw_test.go:84, blindrot/evaluator.go:212)
        assert np.array_equal(ctx.to_host(d)[:, :, : levelQ + 1], want)
    finally:
        ctx.close()


def test_external_product_multiple_p():
    q, p = O.gen_moduli(9, [55, 45, 45, 45, 45], [55, 55])
    _case(8, q, p, 0, 1)
    _case(8, q, p, 0, 2, ct_extra_levels=2)
    _case(8, H.Qi60[:4], H.Pi60[:3], 0, 3)


def test_external_product_single_p_and_bit_decomposition():
    q, p = O.gen_moduli(9, [55, 45, 45], [55])
    _case(8, q, p, 0, 4)                                    # raw (uncentred) limbs as digits: NOT the Decomposer's single-limb rule
    _case(8, q, p, 13, 5)
    _case(8, H.Qi60[:3], H.Pi60[:1], 20, 6)


def test_external_product_without_p_and_32_bit_path():
    q, _ = O.gen_moduli(9, [50, 40], [])
    _case(8, q, [], 10, 7)
    _case(8, q, [], 0, 8)
    q32, _ = O.gen_moduli(11, [27], [])                     # blind-rotation regime: N = 2^10, one modulus below 2^29 (externalProduct32Bit)
    _case(10, q32, [], 7, 9, batch=3)


def test_external_product_fused_sizes():
    q, p = O.gen_moduli(14, [56, 45, 45, 45, 45, 45], [55, 55])
    _case(13, q, p, 0, 10)


@pytest.mark.parametrize("lp,pw2", [([50], 0), ([50, 50], 0), ([], 10)])
def test_blind_rotate_core_matches_the_oracle(lp, pw2):
    """lgpu_blind_rotate_core == oracle/blindrot.py (core/rgsw/blindrot/evaluator.go:144-283) bit for bit: uniform accumulator and keys, masks with
    repeated discrete logs, a zero entry, and window boundaries; the three ExternalProduct paths underneath."""
    import torch
    import lattigo_b200 as lb
    from oracle import blindrot as BR
    logN, n_lwe, window = 8, 24, 4
    q, p = O.gen_moduli(logN + 1, [50, 45], lp)
    params = O.Parameters(logN, q, p)
    N = params.N()
    rng = np.random.default_rng(len(lp) * 7 + pw2)
    levelQ, levelP = len(q) - 1, len(p) - 1
    brk_o = [[H.random_gadget_ciphertext(params, levelQ, levelP, rng, pw2=pw2) for _ in range(2)] for _ in range(n_lwe)]
    gals = [params.GaloisElement(k) for k in range(1, window + 1)] + [2 * N - 5]
    gks_o = {g: H.random_gadget_ciphertext(params, levelQ, levelP, rng, pw2=pw2) for g in gals}
    a = rng.integers(0, N, n_lwe, dtype=np.uint64) * 2 + 1
    a[3] = 0; a[5] = a[7]; a[9] = 1; a[11] = 2 * N - 1
    acc = np.stack([H.rand_poly(q, N, rng), H.rand_poly(q, N, rng)])
    want = [acc[0].copy(), acc[1].copy()]
    BR.Evaluator(params, window).BlindRotateCore(a, want, brk_o, gks_o)
    ctx = lb.Context(logN, q, p)
    try:
        mk = lambda g: lb.GadgetCiphertext(ctx, g.data, levelQ, levelP, pw2, g.pw2_sizes)
        ev = lb.rgsw.BlindRotationEvaluator(ctx, [lb.rgsw.Ciphertext(mk(b[0]), mk(b[1])) for b in brk_o], {g: mk(k) for g, k in gks_o.items()}, window)
        d = ctx.to_device(acc)
        ev.BlindRotateCore(a, d)
        got = ctx.to_host(d)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
        ev2 = lb.rgsw.BlindRotationEvaluator(ctx, ev.brk, {g: k for g, k in ev.keys.items() if g != gals[1]}, window)
        with pytest.raises(lb.LgpuError, match="GaloisKey"):
            ev2.BlindRotateCore(a, ctx.to_device(acc))
    finally:
        ctx.close()
