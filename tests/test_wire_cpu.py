"""Wire-format loaders on the CPU: parsing / sizing / error paths of lgpu_gadget_ct_load with a host-only context (no device
work happens when dst == NULL). The device side is tests/test_gpu_wire.py."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import helpers as H


def _ctx_and_key(pw2=0):
    import lattigo_b200 as lb
    logN = 5
    q, p = O.gen_moduli(logN + 1, [50, 40, 40, 40, 40], [50, 50] if pw2 == 0 else [50])
    params = O.Parameters(logN, q, p)
    rng = np.random.default_rng(3)
    gct = H.random_gadget_ciphertext(params, len(q) - 1, len(p) - 1, rng, pw2=pw2)
    ctx = lb.Context(logN, q, p, device=-1)
    return lb, ctx, params, gct


def test_gadget_ct_info_matches_the_serialised_key():
    lb, ctx, params, gct = _ctx_and_key()
    data = H.marshal_gadget_ct(gct)
    info = lb.wire.gadget_ct_info(ctx, data)
    assert (info.level_q, info.level_p, info.base_two_decomposition) == (4, 1, 0)
    assert info.n_digits == gct.data.shape[0] and info.n_pw2_max == 1
    assert info.consumed == len(data) and info.device_bytes == gct.data.size * 8
    # trailing bytes of a longer stream are not consumed
    assert lb.wire.gadget_ct_info(ctx, data + b"\x01" * 24).consumed == len(data)
    ctx.close()


def test_gadget_ct_info_bit_decomposition_sizes():
    lb, ctx, params, gct = _ctx_and_key(pw2=16)
    info = lb.wire.gadget_ct_info(ctx, H.marshal_gadget_ct(gct))
    assert info.base_two_decomposition == 16 and info.level_p == 0
    assert [info.pw2_sizes[i] for i in range(info.n_digits)] == list(gct.pw2_sizes) and info.n_pw2_max == max(gct.pw2_sizes)
    ctx.close()


def test_malformed_streams_are_errors():
    lb, ctx, params, gct = _ctx_and_key()
    data = H.marshal_gadget_ct(gct)
    for bad, msg in ((data[: len(data) // 2], "truncated"), (data[:12], "truncated"),
                     (data[:16] + np.array([1, 3], dtype="<u8").tobytes() + data[32:], "degree-1")):
        with pytest.raises(lb.LgpuError, match=msg):
            lb.wire.gadget_ct_info(ctx, bad)
    # a row with the wrong number of coefficients (another ring degree)
    words = np.frombuffer(data, dtype="<u8").copy()
    words[5] = 16                                            # first row length of the first polynomial
    with pytest.raises(lb.LgpuError, match="ring degree"):
        lb.wire.gadget_ct_info(ctx, words.tobytes())
    ctx.close()
