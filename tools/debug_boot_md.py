"""Debug helper: fused ModDown (QP-stacked accumulators) vs oracle on the BOOT chain, per row, with details of the bad words."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lattigo_b200 as lb
from lattigo_b200 import params as presets
from oracle import oracle as O
from tests import helpers as H

s = presets.PRESETS["BOOT_N16QP1767"]
q, p = s["Q"], s["P"]
logN = 16
levels = [int(x) for x in os.environ.get("DBG_LEVELS", "23").split(",")]
ctx = lb.Context(logN, q, p)
params = O.Parameters(logN, q, p)
N = params.N()
rng = np.random.default_rng(5)
ev = lb.Evaluator(ctx)
be_o = O.BasisExtender(params.ringQ, params.ringP)
levelP = len(p) - 1
for levelQ in levels:
    nq, npp = levelQ + 1, levelP + 1
    acc = np.stack([np.concatenate([H.rand_poly(q[:nq], N, rng), H.rand_poly(p, N, rng)]) for _ in range(2)])   # [2][nq+np][N]
    want = np.zeros((2, nq, N), dtype=np.uint64)
    for k in range(2):
        be_o.ModDownQPtoQNTT(levelQ, levelP, acc[k, :nq].copy(), acc[k, nq:].copy(), want[k])
    d = ctx.to_device(acc)
    out = torch.zeros((2, nq, N), dtype=torch.int64, device=d.device)
    ev.ModDown(levelQ, levelP, d[0, :nq], d[0, nq:], d[1, :nq], d[1, nq:], out[0], out[1])
    g = ctx.to_host(out)
    for k in range(2):
        for i in range(nq):
            bad = np.nonzero(g[k, i] != want[k, i])[0]
            if len(bad):
                j = int(bad[0]); qi = q[i]
                diff = (int(g[k, i, j]) - int(want[k, i, j])) % qi
                print("level", levelQ, "comp", k, "row", i, "bits", qi.bit_length(), "nbad", len(bad), "idx", j, "chunk", j >> 12, "in-chunk", j & 4095,
                      "got", int(g[k, i, j]), "want", int(want[k, i, j]), "diff mod q", diff, "q-diff", qi - diff)
    print("level", levelQ, "done; any bad:", not np.array_equal(g, want))
