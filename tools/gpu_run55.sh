#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for pr in 256 128 0; do
AUM_TMA_L2PROMO=$pr timeout 200 python tools/kernel_bench.py --only gemm --batch 32 > gpurun_out/kb55_$pr.log 2>&1; echo "promo=$pr rc=$?"; grep -E "\"gemm_(in_proj|out_proj|x_proj)\"" gpurun_out/kb55_$pr.log | cut -c1-100
done
