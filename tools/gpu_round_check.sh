#!/bin/bash
# One gpurun call: GPU parity suite, contract bench, NTT A/B, ncu launch list, ncu --set full of the top kernels.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 600 gpurun_out/bench.json
python tools/bench_ntt.py --out gpurun_out/ntt_ab.json --iters 10 > gpurun_out/ntt_ab.log 2>&1; cat gpurun_out/ntt_ab.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --batch 16 > gpurun_out/launch_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ntt_persist -c 2 -o gpurun_out/ncu_ntt_fwd -f python tools/prof_ntt.py 8 ckks45 fwd > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'ks_chunk_mac_fp8r|ks_strided_j4' -c 4 -o gpurun_out/ncu_ks -f python tools/prof_step.py CKKS_L44 16 > /dev/null 2>&1
ls -la gpurun_out
