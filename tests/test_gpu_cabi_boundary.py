"""The C ABI used from plain C the way INTEGRATION.md's cgo shim uses it (tests/cabi/boundary.c): cudaMallocManaged
polynomials written / read by the host around lgpu_sync, caller tables through lgpu_ring_set_roots, a key switch on managed
memory, and 8 threads x 8 streams hammering two entry points. The program is compiled here with gcc against the public
header and the in-tree library; the non-GPU half of this file checks that it compiles and links as strict C."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cabi", "boundary.c")
LIBDIR = os.path.join(ROOT, "lattigo_b200", "lib")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def _build(tmp_path):
    exe = str(tmp_path / "boundary")
    cmd = ["gcc", "-std=c99", "-O1", "-Wall", "-Werror", SRC, "-I", os.path.join(ROOT, "include"), "-I", os.path.join(CUDA, "include"),
           "-L", LIBDIR, "-llattigo_b200", "-L", os.path.join(CUDA, "lib64"), "-lcudart", "-lpthread",
           "-Wl,-rpath," + LIBDIR, "-Wl,-rpath," + os.path.join(CUDA, "lib64"), "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_boundary_program_compiles_and_links_as_c99(tmp_path):
    if not os.path.exists(os.path.join(LIBDIR, "liblattigo_b200.so")):
        pytest.skip("library not built")
    _build(tmp_path)


@pytest.mark.gpu
def test_boundary_program_runs(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "boundary ok" in r.stdout, (r.stdout[-500:], r.stderr[-2000:])
