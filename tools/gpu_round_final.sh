#!/bin/bash
# Final evidence run of the round (one gpurun call): GPU parity suite, smoke, contract bench (both arms), launch list, ncu captures of the hot kernels,
# config table, lintrans and bootstrap benches. Outputs land in gpurun_out/ and are copied into profiles/ by hand.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_final.log; tail -3 gpurun_out/pytest_gpu_final.log
python __graft_entry__.py smoke > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_final.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 300 gpurun_out/bench_final.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null
LGPU_K3_VARIANT=13 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('LGPU_K3_VARIANT=13', round(d['value'],1), {k: round(v['ms'],1) for k,v in d['roofline']['classes'].items()})"
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --batch 16 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'ks_chunk_mac|ks_strided_j4' -c 5 -o gpurun_out/ncu_ks_final -f python tools/prof_step.py CKKS_L44 16 > /dev/null 2>&1
python tools/bench_configs.py --out gpurun_out/configs_final.json > /dev/null 2>&1
python tools/bench_lintrans.py --out gpurun_out/lintrans_final.json > /dev/null 2>&1
python bench.py --workload bootstrap --preset BOOT_N16QP1767 --batch 64 --steps 2 --warmup 1 > gpurun_out/boot_n1_final.json 2>/dev/null
ls -la gpurun_out | tail -12
