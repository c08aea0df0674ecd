"""Standalone transform benchmark (BASELINE config 2: N = 2^16, 44 limbs) for A/B of kernel variants. Each variant is a
set of LGPU_* environment switches and runs in its own process (the switches are read once per process).

  python tools/bench_ntt.py --out gpurun_out/ntt_ab.json            all variants
  python tools/bench_ntt.py --worker                                one variant (internal)

Timing: CUDA events on the launching stream, median of `iters`, L2 flushed (256 MiB write) before every iteration.
Algorithmic bytes = 16 B per coefficient (read once + write once; twiddle tables excluded).
Also checks INTT(NTT(x)) == x and NTT linearity on the device as a cheap guard (parity proper is tests/)."""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

VARIANTS = {
    "two_pass": {"LGPU_NTT_PERSIST": "0"},                                 # round-1 kernels: strided pass + chunk pass, 2 HBM round trips
    "persist_v1": {"LGPU_NTT_PERSIST": "1", "LGPU_NTT_PERSIST_V": "1"},    # single HBM pass, padded tile, CTA barriers, serial ticket
    "persist_v2": {"LGPU_NTT_PERSIST": "1", "LGPU_NTT_PERSIST_V": "2"},    # + claim-ahead tile loop
    "persist_v3": {"LGPU_NTT_PERSIST": "1", "LGPU_NTT_PERSIST_V": "3"},    # + swizzled tile, pair / warp level exchanges, hoisted addresses (default)
    "persist_v4": {"LGPU_NTT_PERSIST": "1", "LGPU_NTT_PERSIST_V": "4"},    # v3 + TMA bulk prefetch of the next chunk into a double-buffered tile
    "persist_v5": {"LGPU_NTT_PERSIST": "1", "LGPU_NTT_PERSIST_V": "5"},    # v3 + register copy-out (128-bit stores straight from the last round)
    "persist_v6": {"LGPU_NTT_PERSIST": "1", "LGPU_NTT_PERSIST_V": "6"},    # v5 under a static tile schedule with a double-buffered tile (one CTA barrier per tile)
    "persist_ph1int": {"LGPU_NTT_PERSIST": "1", "LGPU_NTT_PH1INT": "1"},   # strided stages on the integer pipes (v1 tile code)
    "int_256x16": {"LGPU_NTT_PERSIST_IV": "2"},                            # integer rows: 256 x 16 padded tile (IntFwdOps) instead of the default 512 x 8 swizzled one
}


def worker(args):
    import numpy as np
    import torch
    import lattigo_b200 as lb
    from lattigo_b200 import params as presets
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(7)
    flush = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
    N, nl = 1 << 16, 44
    peak = 6582.5
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    res = {}
    for tag, Q in (("ckks45", presets.PRESETS["CKKS_L44"]["Q"][:nl]), ("q61", presets.QI60[:32] + presets.PI60[:12])):
        ctx = lb.Context(16, Q)
        for batch in (1, 8, 16):
            x = torch.empty((batch, nl, N), dtype=torch.int64, device=dev)
            for i, q in enumerate(Q):
                x[:, i] = torch.from_numpy(rng.integers(0, q, size=(batch, N), dtype=np.uint64).view(np.int64)).to(dev)
            y = torch.empty_like(x); z = torch.empty_like(x)
            rq = ctx.ringQ
            rq.NTT(x, y); rq.INTT(y, z)
            ok = bool(torch.equal(x, z))
            for name, fn in (("ntt", lambda: rq.NTT(x, y)), ("intt", lambda: rq.INTT(y, z))):
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                ts = []
                for _ in range(args.iters):
                    flush.add_(1)
                    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
                    a.record(); fn(); b.record(); torch.cuda.synchronize()
                    ts.append(a.elapsed_time(b) * 1e-3)
                t = float(np.median(ts))
                alg = 16.0 * N * nl * batch
                res["%s_b%d_%s" % (tag, batch, name)] = {"us": t * 1e6, "us_per_limb": t * 1e6 / (nl * batch), "alg_GBs": alg / t / 1e9,
                                                         "frac_of_measured_hbm": alg / t / 1e9 / peak, "roundtrip_ok": ok}
        ctx.close()
    print(json.dumps(res))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/ntt_ab.json")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--worker", action="store_true")
    ap.add_argument("--variants", default="")
    ap.add_argument("--extra", default="", help="extra variants: name:K=V,K=V;name2:...")
    args = ap.parse_args()
    if args.worker:
        return worker(args)
    variants = dict(VARIANTS)
    for spec in filter(None, args.extra.split(";")):
        name, kv = spec.split(":")
        variants[name] = dict(x.split("=") for x in kv.split(","))
    names = (args.variants.split(",") if args.variants else list(VARIANTS)) + [n for n in variants if n not in VARIANTS]
    out = {}
    for v in names:
        env = dict(os.environ, **variants[v])
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", "--iters", str(args.iters)], env=env,
                           capture_output=True, text=True, timeout=900)
        try:
            out[v] = {"env": variants[v], "results": json.loads(r.stdout.strip().splitlines()[-1])}
        except Exception:
            out[v] = {"env": variants[v], "error": (r.stdout + r.stderr)[-2000:]}
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    json.dump(out, open(args.out, "w"), indent=1)
    for v, d in out.items():
        if "results" in d:
            print(v, {k: round(x["frac_of_measured_hbm"], 3) for k, x in d["results"].items()},
                  "ok" if all(x["roundtrip_ok"] for x in d["results"].values()) else "ROUNDTRIP MISMATCH")
        else:
            print(v, "ERROR", d["error"][-500:])


if __name__ == "__main__":
    main()
