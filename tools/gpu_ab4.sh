#!/bin/bash
# Last validation of the round: full GPU suite + smoke + contract bench on the final tree, then A/B of the 256-bit key / operand accesses
# (LGPU_K3_WIDE=0 switches them off) and the bootstrap replay (clock poller now started before the warm-up).
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_ab4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_ab4.log; tail -3 gpurun_out/pytest_gpu_ab4.log
python __graft_entry__.py smoke > gpurun_out/smoke_ab4.log 2>&1; echo "smoke rc=$?"
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_ab4.json 2> gpurun_out/bench_ab4.err; tail -c 150 gpurun_out/bench_ab4.json; echo
for cfg in "LGPU_X=0" "LGPU_K3_WIDE=0"; do
  env $cfg timeout 200 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$cfg', round(d['value'],1), round(d['ms_per_step'],2), {k: round(v['ms'],1) for k,v in d['roofline']['classes'].items()})"
done
timeout 200 python bench.py --workload bootstrap --preset BOOT_N16QP1767 --batch 64 --steps 2 --warmup 1 > gpurun_out/boot_ab4.json 2>/dev/null; python -c "
import json
d=json.loads(open('gpurun_out/boot_ab4.json').read().strip().splitlines()[-1]); print('bootstrap', round(d['value'],2), round(d['ms_per_step'],1), d['phase_ms_per_step'], d['clocks'])"
