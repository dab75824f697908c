#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 83 -c 1 -o gpurun_out/gemm_dtsp_v1 -f python tools/kernel_bench.py --only gemm --batch 32 > gpurun_out/ncu_dtsp.log 2>&1; echo "ncu rc=$?"
