#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -x -k "gemm" > gpurun_out/t_g.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_g.log | cut -c1-200
for b in 64 32; do
timeout 200 python tools/kernel_bench.py --only gemm --batch $b > gpurun_out/kb57_$b.log 2>&1; echo "batch=$b rc=$?"; grep -E "gemm_" gpurun_out/kb57_$b.log | cut -c1-105
done
AUM_GEMM_PAIR=0 timeout 200 python tools/kernel_bench.py --only gemm --batch 64 > gpurun_out/kb57_np.log 2>&1; echo "nopair batch=64 rc=$?"; grep -E "\"gemm_(in_proj|out_proj)\"" gpurun_out/kb57_np.log | cut -c1-105
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench57.json 2> gpurun_out/bench57.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench57.json')); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'])"
