#!/bin/bash
# two-rank runs of both bench workloads (weak scaling, one process per GPU)
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload bootstrap --preset BOOT_N16QP1767 --batch 64 --steps 2 --warmup 1 > gpurun_out/boot_2gpu.json 2> gpurun_out/boot_2gpu.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_2gpu.json", "gpurun_out/boot_2gpu.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], round(d["value"], 1), d["unit"], round(d["ms_per_step"], 1), (d.get("e2e") or {}).get("value"))
    except Exception as ex:
        print(f, "FAILED", ex); print(open(f.replace(".json", ".err")).read()[-1500:])
PY
