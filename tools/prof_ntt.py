"""Tiny driver for ncu captures of the NTT kernels at the C2 shape (N=2^16, 44 limbs)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lattigo_b200 as lb  # noqa: E402
from lattigo_b200 import params as presets  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 2
which = sys.argv[2] if len(sys.argv) > 2 else "q61"
Q = presets.QI60[:32] + presets.PI60[:12] if which == "q61" else presets.PRESETS["CKKS_L44"]["Q"][1:] + presets.PRESETS["CKKS_L44"]["Q"][:1]
ctx = lb.Context(16, Q)
x = torch.randint(0, 2**60, (batch, 44, 1 << 16), dtype=torch.int64, device="cuda")
y = torch.empty_like(x)
for _ in range(2):
    ctx.ringQ.NTT(x, y)
    ctx.ringQ.INTT(x, y)
torch.cuda.synchronize()
