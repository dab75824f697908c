#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
AUM_GEMM_LITE=1 timeout 300 python tools/kernel_bench.py --only gemm > gpurun_out/kb33.log 2>&1; echo "kb rc=$?"; grep gemm_ gpurun_out/kb33.log | cut -c1-120
for lite in 0 1; do
AUM_GEMM_LITE=$lite timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench33_$lite.json 2> gpurun_out/bench33.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench33_$lite.json')); print('lite=$lite', {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'])"
done
