#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -x -k "gemm" > gpurun_out/t_g.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_g.log | cut -c1-200
for b in 64 32; do
timeout 200 python tools/kernel_bench.py --only gemm --batch $b > gpurun_out/kb61_$b.log 2>&1; echo "batch=$b rc=$?"; grep -E "dt_proj" gpurun_out/kb61_$b.log | cut -c1-105
done
