"""GPU side of the wire loaders (csrc/wire.cu): a key serialised in the reference's format
(core/rlwe/gadgetciphertext.go:101-121) and loaded with lgpu_gadget_ct_load / lgpu_galois_key_load key-switches exactly like
the same key uploaded as a raw tensor and like the oracle; ring.Poly round-trips through lgpu_poly_load / lgpu_poly_store."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu
U64 = np.uint64


@pytest.mark.parametrize("pw2", [0, 16])
def test_gadget_product_through_a_wire_loaded_key(pw2):
    import lattigo_b200 as lb
    logN = 9
    q, p = O.gen_moduli(logN + 1, [55, 45, 45, 45, 45], [55, 55] if pw2 == 0 else [55])
    params = O.Parameters(logN, q, p)
    N = params.N()
    rng = np.random.default_rng(4)
    level, levelP = len(q) - 1, len(p) - 1
    gct = H.random_gadget_ciphertext(params, level, levelP, rng, pw2=pw2)
    ctx = lb.Context(logN, q, p)
    try:
        raw = H.marshal_galois_key(params.GaloisElement(3), 2 * N, gct)
        g, key = lb.wire.load_galois_key(ctx, raw)
        assert g == params.GaloisElement(3)
        assert (key.levelQ, key.levelP, key.BaseTwoDecomposition) == (level, levelP, pw2)
        ref_key = lb.GadgetCiphertext(ctx, gct.data, level, levelP, pw2, gct.pw2_sizes)
        assert np.array_equal(ctx.to_host(key.data), ctx.to_host(ref_key.data))
        cx = np.stack([H.rand_poly(q, N, rng) for _ in range(2)])
        ev = lb.Evaluator(ctx)
        d0 = ctx.new_poly(level + 1, 2); d1 = ctx.new_poly(level + 1, 2)
        ev.GadgetProduct(level, ctx.to_device(cx), key, d0, d1)
        ev_o = O.Evaluator(params)
        for b in range(2):
            w0 = np.zeros((level + 1, N), dtype=U64); w1 = np.zeros((level + 1, N), dtype=U64)
            ev_o.GadgetProduct(level, cx[b], gct, [w0, w1])
            assert np.array_equal(ctx.to_host(d0)[b], w0) and np.array_equal(ctx.to_host(d1)[b], w1)
        with pytest.raises(lb.LgpuError, match="NthRoot"):
            lb.wire.load_galois_key(ctx, H.marshal_galois_key(5, 4 * N, gct))
    finally:
        ctx.close()


def test_poly_load_store_round_trip():
    import lattigo_b200 as lb
    logN = 10
    q, _ = O.gen_moduli(logN + 1, [55, 45, 45], [])
    ctx = lb.Context(logN, q)
    try:
        rng = np.random.default_rng(9)
        x = H.rand_poly(q, 1 << logN, rng)
        raw = H.marshal_poly(x)
        t, used = lb.wire.load_poly(ctx, raw + b"\x00" * 8, 3)
        assert used == len(raw) and np.array_equal(ctx.to_host(t), x)
        assert lb.wire.store_poly(ctx, t) == raw
        with pytest.raises(lb.LgpuError, match="fewer rows"):
            lb.wire.load_poly(ctx, raw, 2)
    finally:
        ctx.close()
