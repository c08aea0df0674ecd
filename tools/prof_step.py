"""One small CKKS MulRelin+Rescale step (8 ciphertext pairs at the bench preset) for ncu captures."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lattigo_b200 as lb  # noqa: E402
from lattigo_b200 import params as presets  # noqa: E402

s = presets.PRESETS[sys.argv[1] if len(sys.argv) > 1 else "CKKS_L44"]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
logN, Q, P = s["logN"], s["Q"], s["P"]
N = 1 << logN
ctx = lb.Context(logN, Q, P)
g = torch.Generator(device="cuda"); g.manual_seed(1)


def rand_rows(mods, lead):
    out = torch.empty(tuple(lead) + (len(mods), N), dtype=torch.int64, device="cuda")
    for i, m in enumerate(mods):
        out[..., i, :] = torch.randint(0, m, tuple(lead) + (N,), generator=g, device="cuda", dtype=torch.int64)
    return out


level, levelP = len(Q) - 1, len(P) - 1
nd = (level + levelP + 1) // (levelP + 1)
rlk = lb.GadgetCiphertext(ctx, rand_rows(Q + P, (nd, 1, 2)), level, levelP)
ev = lb.CKKSEvaluator(ctx, rlk)
a, b = rand_rows(Q, (B, 2)), rand_rows(Q, (B, 2))
for _ in range(2):
    out = ev.MulRelinRescaleNew(a, b)
torch.cuda.synchronize()
