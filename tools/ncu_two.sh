#!/bin/bash
# ncu --set full of dt_proj (16-bit delta) and the three-term conv backward at their in-model sizes
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
B=32 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 2 -c 1 -o gpurun_out/r2_dt_proj16 -f python tools/one_kernel.py dt_proj16 > gpurun_out/ncu_dt16.log 2>&1; echo "dt16 rc=$?"
B=32 DT=bf16 timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv1d_bwd -s 2 -c 1 -o gpurun_out/r2_conv1d_bwd3 -f python tools/one_kernel.py conv1d_bwd > gpurun_out/ncu_cb.log 2>&1; echo "cb rc=$?"
