#!/bin/bash
# Validation + A/B of the vectorised twiddle loads (64/128/256-bit runs, issued in the digit loop instead of hoisted-and-spilled): full GPU suite,
# smoke, headline bench under K3 variants 11 (default) / 13 / 10, standalone NTT, contract bench, one ncu capture of the MAC kernel.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_ab3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_ab3.log; tail -3 gpurun_out/pytest_gpu_ab3.log
python __graft_entry__.py smoke > gpurun_out/smoke_ab3.log 2>&1; echo "smoke rc=$?"
for cfg in "LGPU_X=0" "LGPU_K3_VARIANT=13" "LGPU_K3_VARIANT=10"; do
  env $cfg timeout 300 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$cfg', round(d['value'],1), round(d['ms_per_step'],2), {k: round(v['ms'],1) for k,v in d['roofline']['classes'].items()}, d['roofline']['ntt_standalone']['us_per_limb_transform'])"
done
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_ab3.json 2> gpurun_out/bench_ab3.err; tail -c 200 gpurun_out/bench_ab3.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'ks_chunk_mac_fp8r|fz_chunk_epi_fp8' -c 2 -o gpurun_out/ncu_k3_ab3 -f python tools/prof_step.py CKKS_L44 16 > /dev/null 2>&1
ls -la gpurun_out | grep ab3
