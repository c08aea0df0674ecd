"""Kernel-level micro-benchmarks (CUDA events on the launching stream, warm-up, L2 flush between iterations).
Not the contract bench (that is bench.py); used to steer kernel work. Prints JSON lines."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lattigo_b200 as lb  # noqa: E402
from lattigo_b200 import params as presets  # noqa: E402


def timeit(fn, iters=20, warmup=3, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.add_(1)
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    return float(np.median(ts)), float(np.min(ts))


def main():
    logN, nl = 16, 44
    N = 1 << logN
    which = sys.argv[1] if len(sys.argv) > 1 else "q61"
    Q = presets.QI60[:32] + presets.PI60[:12] if which == "q61" else presets.PRESETS["CKKS_L44"]["Q"][1:] + [presets.PRESETS["CKKS_L44"]["Q"][1] - 0] [:0] + presets.PRESETS["CKKS_L44"]["Q"][:1]
    Q = Q[:nl]
    ctx = lb.Context(logN, Q)
    rq = ctx.ringQ
    flush = torch.zeros(256 << 20, dtype=torch.uint8, device="cuda")   # 256 MiB > L2
    S = nl * N * 8
    print(json.dumps({"gpu": torch.cuda.get_device_name(0), "cpus": os.cpu_count(), "primes": which}))
    for batch in (1, 4, 8):
        x = torch.randint(0, 2**60, (batch, nl, N), dtype=torch.int64, device="cuda")
        y = torch.empty_like(x)
        z = torch.randint(0, 2**60, (batch, nl, N), dtype=torch.int64, device="cuda")
        for name, fn, bytes_alg in (
            ("ntt_fwd", lambda: rq.NTT(x, y), 2 * S * batch),
            ("ntt_fwd_lazy", lambda: rq.NTTLazy(x, y), 2 * S * batch),
            ("ntt_inv", lambda: rq.INTT(x, y), 2 * S * batch),
            ("mulcoeffs_montgomery", lambda: rq.MulCoeffsMontgomery(x, z, y), 3 * S * batch),
        ):
            med, best = timeit(fn, flush=flush)
            print(json.dumps({"kernel": name, "batch": batch, "limbs": nl, "logN": logN, "us_median": med * 1e6,
                              "us_best": best * 1e6, "alg_GBs_median": bytes_alg / med / 1e9,
                              "us_per_limb": med * 1e6 / (nl * batch)}))


if __name__ == "__main__":
    main()
