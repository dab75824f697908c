#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_mixer_gpu.py tests/test_model_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/t_k.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|Error" gpurun_out/t_k.log | tail -8
for b in 64 32; do
timeout 300 python tools/kernel_bench.py --only gemm --batch $b > gpurun_out/kb42_$b.log 2>&1; echo "kb batch=$b rc=$?"; grep -E "gemm_" gpurun_out/kb42_$b.log | cut -c1-110
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench42.json 2> gpurun_out/bench42.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench42.json')); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'])"
