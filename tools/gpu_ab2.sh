#!/bin/bash
# A/B of the late round-2 variants in ONE gpurun call: NTT copy-out / static-schedule loop (LGPU_NTT_PERSIST_V=5,6), 23-bit-halves basis
# extension in K2 (LGPU_K2_SPLIT=1), FP64-pipe MAC in K3 (LGPU_K3_VARIANT=11,12). Parity first (GPU suite slices under each switch), then timing.
mkdir -p gpurun_out
for v in 5 6; do
  LGPU_NTT_PERSIST=1 LGPU_NTT_PERSIST_V=$v timeout 400 python -m pytest -x -q -m gpu tests/test_gpu_ring.py tests/test_gpu_headline_parity.py > gpurun_out/ab2_ntt_v$v.log 2>&1
  echo "NTT v$v parity: $(tail -1 gpurun_out/ab2_ntt_v$v.log)"
done
timeout 400 python tools/bench_ntt.py --variants persist_v3,persist_v5,persist_v6 --iters 8 --out gpurun_out/ntt_ab2.json > /dev/null 2>&1
python - <<'PY'
import json
d = json.load(open("gpurun_out/ntt_ab2.json"))
for v, r in d.items():
    if "results" in r:
        print(v, {k: round(x["us_per_limb"], 3) for k, x in r["results"].items() if "b16" in k or "b1_" in k}, all(x["roundtrip_ok"] for x in r["results"].values()))
    else:
        print(v, "ERROR", r.get("error", "")[-300:])
PY
for cfg in "LGPU_K2_SPLIT=1" "LGPU_K3_VARIANT=11" "LGPU_K3_VARIANT=12"; do
  tag=$(echo $cfg | tr '=' '_')
  env $cfg timeout 500 python -m pytest -x -q -m gpu tests/test_gpu_headline_parity.py tests/test_gpu_keyswitch.py tests/test_gpu_lintrans.py tests/test_gpu_rgsw.py tests/test_gpu_bootreplay.py > gpurun_out/ab2_$tag.log 2>&1
  echo "$cfg parity: $(tail -1 gpurun_out/ab2_$tag.log)"
done
for cfg in "LGPU_X=0" "LGPU_K2_SPLIT=1" "LGPU_K3_VARIANT=11" "LGPU_K3_VARIANT=12" "LGPU_K2_SPLIT=1 LGPU_K3_VARIANT=11" "LGPU_K2_SPLIT=1 LGPU_K3_VARIANT=12"; do
  tag=$(echo $cfg | tr '= ' '__')
  env $cfg timeout 300 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ab2_bench_$tag.json 2> gpurun_out/ab2_bench_$tag.err
  python - "$cfg" gpurun_out/ab2_bench_$tag.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ct/s", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "K3 frac", round(d["roofline"]["frac"], 4), "classes", {k: round(v["ms"], 1) for k, v in d["roofline"]["classes"].items()})
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
