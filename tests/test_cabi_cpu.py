"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/lattigo_b200.h declares, and its host-side table generation (device < 0: no CUDA call) matches the
oracle's independent big-integer generation."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import helpers as H


def test_library_exports_every_declared_symbol():
    import lattigo_b200 as lb
    L = lb.lib()
    syms = lb.declared_symbols()
    assert len(syms) >= 15
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    assert b"sm_100a" in L.lgpu_version()


def test_host_tables_match_oracle():
    import lattigo_b200 as lb
    logN, N = 11, 2048
    Q, P = H.Qi60[:4], H.Pi60[:3]
    ctx = lb.Context(logN, Q, P, device=-1)
    for ring, mods in ((0, Q), (1, P)):
        oring = O.Ring(N, mods)
        for i, q in enumerate(mods):
            s = oring.SubRings[i]
            t = [int(x) for x in ctx.table(ring, i, 0)]
            assert t == [q, s.MRedConstant, s.BRedConstant[0], s.BRedConstant[1], s.NInv, s.PrimitiveRoot]
            assert np.array_equal(ctx.table(ring, i, 1), s.RootsForward)
            assert np.array_equal(ctx.table(ring, i, 2), s.RootsBackward)
            if i >= 1:
                assert [int(x) for x in ctx.table(ring, i, 3)] == oring.RescaleConstants[i - 1]
    ctx.close()


def test_host_only_context_refuses_device_work_and_bad_params():
    import lattigo_b200 as lb
    ctx = lb.Context(8, H.Qi60[:2], device=-1)
    with pytest.raises(lb.LgpuError):
        ctx.sync()
    with pytest.raises(lb.LgpuError):
        lb.Context(3, H.Qi60[:1], device=-1)          # N < 16
    with pytest.raises(lb.LgpuError):
        lb.Context(8, [H.Qi60[0], H.Qi60[0]], device=-1)
    with pytest.raises(lb.LgpuError):
        lb.Context(8, [97], device=-1)                 # 97 != 1 mod 512
    ctx.close()


def test_header_is_plain_c():
    """The boundary is a C ABI: include/lattigo_b200.h must compile as C99 (what cgo feeds to the C compiler) and as C++,
    with plain pointers / sizes only (no torch or CUDA types in any signature)."""
    import os
    import shutil
    import subprocess
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    inc = os.path.join(root, "include")
    import re
    src = re.sub(r"/\*.*?\*/", "", open(os.path.join(inc, "lattigo_b200.h")).read(), flags=re.S)    # declarations only
    for banned in ("torch", "at::", "cudaStream_t", "std::", "#include <cuda"):
        assert banned not in src, banned
    gcc = shutil.which("gcc") or "/opt/gcc/bin/gcc"
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "h.c")
        open(c, "w").write('#include "lattigo_b200.h"\nint main(void) { return (int)LGPU_OP_COUNT - (int)LGPU_OP_COUNT; }\n')
        r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", inc, c], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        gpp = shutil.which("g++")
        if gpp:
            r = subprocess.run([gpp, "-std=c++17", "-fsyntax-only", "-x", "c++", "-I", inc, c], capture_output=True, text=True)
            assert r.returncode == 0, r.stderr
