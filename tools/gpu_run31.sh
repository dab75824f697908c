#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv1d_fwd -s 5 -c 1 -o gpurun_out/conv_v1 -f python tools/kernel_bench.py --only conv > gpurun_out/ncu_conv.log 2>&1; echo "ncu conv rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:add_rmsnorm -s 5 -c 1 -o gpurun_out/norm_v1 -f python tools/kernel_bench.py --only norm > gpurun_out/ncu_norm.log 2>&1; echo "ncu norm rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 57 -c 1 -o gpurun_out/gemm_dt_v1 -f python tools/kernel_bench.py --only gemm > gpurun_out/ncu_dt.log 2>&1; echo "ncu dt rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 44 -c 1 -o gpurun_out/gemm_x_v1 -f python tools/kernel_bench.py --only gemm > gpurun_out/ncu_x.log 2>&1; echo "ncu x rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 31 -c 1 -o gpurun_out/gemm_out_v1 -f python tools/kernel_bench.py --only gemm > gpurun_out/ncu_out.log 2>&1; echo "ncu out rc=$?"
