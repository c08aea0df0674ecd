"""Tiny driver for ncu captures of the transform kernels at the C2 shape (N=2^16, 44 limbs).
usage: prof_ntt.py <batch> <q61|ckks45> [fwd|inv|both]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lattigo_b200 as lb  # noqa: E402
from lattigo_b200 import params as presets  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 2
which = sys.argv[2] if len(sys.argv) > 2 else "q61"
what = sys.argv[3] if len(sys.argv) > 3 else "both"
Q = presets.QI60[:32] + presets.PI60[:12] if which == "q61" else presets.PRESETS["CKKS_L44"]["Q"][1:] + presets.PRESETS["CKKS_L44"]["Q"][:1]
ctx = lb.Context(16, Q)
x = torch.empty((batch, 44, 1 << 16), dtype=torch.int64, device="cuda")
for i, q in enumerate(Q):
    x[:, i] = torch.randint(0, q, (batch, 1 << 16), dtype=torch.int64, device="cuda")
y = torch.empty_like(x)
for _ in range(2):
    if what in ("fwd", "both"):
        ctx.ringQ.NTT(x, y)
    if what in ("inv", "both"):
        ctx.ringQ.INTT(x, y)
torch.cuda.synchronize()
