#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for b in 64 32; do
timeout 300 python tools/kernel_bench.py --only gemm --batch $b > gpurun_out/kb40_$b.log 2>&1; echo "kb batch=$b rc=$?"; grep -E "x_proj|dt_proj" gpurun_out/kb40_$b.log | cut -c1-120
done
AUM_GEMM_DIRECT_STORE=1 timeout 300 python tools/kernel_bench.py --only gemm --batch 32 > gpurun_out/kb40_direct.log 2>&1; echo "kb direct rc=$?"; grep -E "x_proj|dt_proj" gpurun_out/kb40_direct.log | cut -c1-120
