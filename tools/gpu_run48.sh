#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for bn in 96 128 192 256; do for b in 32 64; do
AUM_GEMM_BN=$bn timeout 300 python tools/kernel_bench.py --only gemm --batch $b > gpurun_out/kb48_${bn}_$b.log 2>&1; echo "bn=$bn batch=$b rc=$?"; grep -E "\"gemm_in_proj\"|\"gemm_out_proj\"" gpurun_out/kb48_${bn}_$b.log | cut -c1-100
done; done
