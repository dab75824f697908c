#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for h in 0 1; do
if [ $h = 1 ]; then export AUM_GEMM_HACK_NACC=1; fi
AUM_GEMM_PAIR=0 timeout 200 python tools/kernel_bench.py --only gemm --batch 32 > gpurun_out/kb56_$h.log 2>&1; echo "hack=$h rc=$?"; grep -E "\"gemm_(in_proj|out_proj|x_proj|dt_proj)\"" gpurun_out/kb56_$h.log | cut -c1-100
done
