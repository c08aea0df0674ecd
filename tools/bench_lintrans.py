"""Throughput of the device linear-transformation driver beside the un-hoisted alternative:
  bsgs      lgpu_lintrans_evaluate_many, N1 = 8 (double hoisting: one decomposition, 7 hoisted baby rotations, 3 giant-step key switches)
  naive     lgpu_lintrans_evaluate_many, N1 = 0 (single hoisting: one decomposition, one hoisted key switch per diagonal)
  unhoisted per diagonal: Evaluator.Automorphism (a full key switch with its own decomposition) then MulCoeffsMontgomeryThenAdd
on a 32-diagonal matrix, CKKS PN16QP1761 (N = 2^16, 34 + 4 limbs) at the top level, batch of 8 ciphertexts, synthetic operands.
  python tools/bench_lintrans.py --out gpurun_out/lintrans.json"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lattigo_b200 as lb  # noqa: E402
from lattigo_b200 import params as presets  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/lintrans.json")
    ap.add_argument("--preset", default="CKKS_PN16QP1761")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--diags", type=int, default=32)
    ap.add_argument("--n1", type=int, default=8)
    ap.add_argument("--iters", type=int, default=3)
    args = ap.parse_args()
    P = presets.PRESETS[args.preset]
    logN, Q, Pp = P["logN"], P["Q"], P["P"]
    N = 1 << logN
    ctx = lb.Context(logN, Q, Pp)
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev); g.manual_seed(3)
    level, levelP = len(Q) - 1, len(Pp) - 1
    nd = (level + levelP + 1) // (levelP + 1)

    def rand_rows(mods, lead):
        out = torch.empty(tuple(lead) + (len(mods), N), dtype=torch.int64, device=dev)
        for i, m in enumerate(mods):
            out[..., i, :] = torch.randint(0, m, tuple(lead) + (N,), generator=g, device=dev, dtype=torch.int64)
        return out

    rots = list(range(args.diags))
    ev = lb.lintrans.Evaluator(ctx, {})
    need = sorted(set(r for r in rots if r) | {(r // args.n1) * args.n1 for r in rots if (r // args.n1) * args.n1} | {r % args.n1 for r in rots if r % args.n1})
    keys = {ev.GaloisElement(r): lb.GadgetCiphertext(ctx, rand_rows(Q + Pp, (nd, 1, 2)), level, levelP) for r in need}
    ev = lb.lintrans.Evaluator(ctx, keys)
    vec = {r: rand_rows(Q + Pp, ()) for r in rots}
    ct = rand_rows(Q, (args.batch, 2))
    out = torch.empty_like(ct)
    res = {"preset": args.preset, "batch": args.batch, "diagonals": args.diags, "N1": args.n1, "galois_keys": len(keys),
           "timing": "CUDA events around %d calls after one warm-up call" % args.iters}

    def timeit(fn):
        fn(); torch.cuda.synchronize()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.iters):
            fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) * 1e-3 / args.iters

    lt_bsgs = lb.lintrans.LinearTransformation(vec, level, levelP, logN - 1, args.n1)
    lt_naive = lb.lintrans.LinearTransformation(vec, level, levelP, logN - 1, 0)
    t = timeit(lambda: ev.Evaluate(ct, lt_bsgs, out))
    res["bsgs"] = {"ms_per_batch": t * 1e3, "ct_per_s": args.batch / t}
    ref_bsgs = out.clone()
    t = timeit(lambda: ev.Evaluate(ct, lt_naive, out))
    res["naive_hoisted"] = {"ms_per_batch": t * 1e3, "ct_per_s": args.batch / t}
    rl = lb.Evaluator(ctx)
    rq = ctx.ringQ
    tmp = torch.empty_like(ct)
    pt_q = {r: vec[r][: level + 1].contiguous() for r in rots}

    def unhoisted():
        for r in rots:
            src = ct
            if r:
                rl.Automorphism(ct, ev.GaloisElement(r), keys[ev.GaloisElement(r)], tmp)
                src = tmp
            for k in range(2):
                for b in range(args.batch):
                    if r == rots[0]:
                        rq.MulCoeffsMontgomery(pt_q[r], src[b, k], out[b, k])
                    else:
                        rq.MulCoeffsMontgomeryThenAdd(pt_q[r], src[b, k], out[b, k])
    t = timeit(unhoisted)
    res["unhoisted_rotate_and_add"] = {"ms_per_batch": t * 1e3, "ct_per_s": args.batch / t}
    res["speedup_bsgs_vs_unhoisted"] = res["bsgs"]["ct_per_s"] / res["unhoisted_rotate_and_add"]["ct_per_s"]
    res["speedup_naive_vs_unhoisted"] = res["naive_hoisted"]["ct_per_s"] / res["unhoisted_rotate_and_add"]["ct_per_s"]
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
