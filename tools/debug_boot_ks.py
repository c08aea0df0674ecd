"""Debug helper: per-row comparison of GadgetProductLazy / GadgetProduct against the oracle on the BOOT chain."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lattigo_b200 as lb
from lattigo_b200 import params as presets
from oracle import oracle as O
from tests import helpers as H

s = presets.PRESETS["BOOT_N16QP1767"]
logN = int(os.environ.get("DBG_LOGN", "16"))
q, p = s["Q"], s["P"]
if logN != 16:
    q, p = O.gen_moduli(logN + 1, s["LogQ"], s["LogP"])
levels = [int(x) for x in os.environ.get("DBG_LEVELS", "20").split(",")]
ctx = lb.Context(logN, q, p)
params = O.Parameters(logN, q, p)
N = params.N()
rng = np.random.default_rng(81)
ev_o = O.Evaluator(params); ev = lb.Evaluator(ctx)
evk_o = H.random_gadget_ciphertext(params, params.MaxLevelQ(), params.MaxLevelP(), rng)
evk = lb.GadgetCiphertext(ctx, evk_o.data, evk_o.LevelQ(), evk_o.LevelP(), 0, evk_o.pw2_sizes)
levelP = params.MaxLevelP()
for levelQ in levels:
    cx = H.rand_poly(q[: levelQ + 1], N, rng)
    wq = [np.zeros((levelQ + 1, N), dtype=np.uint64) for _ in range(2)]; wp = [np.zeros((levelP + 1, N), dtype=np.uint64) for _ in range(2)]
    ev_o.GadgetProductLazy(levelQ, cx.copy(), evk_o, wq, wp)
    want = [np.zeros((levelQ + 1, N), dtype=np.uint64) for _ in range(2)]
    ev_o.GadgetProduct(levelQ, cx.copy(), evk_o, want)
    d = ctx.to_device(cx)
    a0q = ctx.new_poly(levelQ + 1); a1q = ctx.new_poly(levelQ + 1); a0p = ctx.new_poly(levelP + 1); a1p = ctx.new_poly(levelP + 1)
    ev.GadgetProductLazy(levelQ, d, evk, a0q, a0p, a1q, a1p)
    for name, got, w in (("acc0Q", a0q, wq[0]), ("acc1Q", a1q, wq[1]), ("acc0P", a0p, wp[0]), ("acc1P", a1p, wp[1])):
        g = ctx.to_host(got)
        bad = [(i, int((g[i] != w[i]).sum())) for i in range(g.shape[0]) if not np.array_equal(g[i], w[i])]
        print("level", levelQ, name, "bad rows:", bad)
    c0 = ctx.new_poly(levelQ + 1); c1 = ctx.new_poly(levelQ + 1)
    ev.GadgetProduct(levelQ, d, evk, c0, c1)
    for name, got, w in (("ct0", c0, want[0]), ("ct1", c1, want[1])):
        g = ctx.to_host(got)
        bad = [(i, int((g[i] != w[i]).sum())) for i in range(g.shape[0]) if not np.array_equal(g[i], w[i])]
        print("level", levelQ, name, "bad rows:", bad)
    m0 = ctx.new_poly(levelQ + 1); m1 = ctx.new_poly(levelQ + 1)
    ev.ModDown(levelQ, levelP, ctx.to_device(wq[0]), ctx.to_device(wp[0]), ctx.to_device(wq[1]), ctx.to_device(wp[1]), m0, m1)
    for name, got, w in (("moddown0", m0, want[0]), ("moddown1", m1, want[1])):
        g = ctx.to_host(got)
        bad = [(i, int((g[i] != w[i]).sum())) for i in range(g.shape[0]) if not np.array_equal(g[i], w[i])]
        print("level", levelQ, name, "bad rows (oracle accumulators in):", bad)
