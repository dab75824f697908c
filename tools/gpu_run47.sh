#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python tools/kernel_bench.py --only bwd --batch 32 --dtype bf16 > gpurun_out/kb47.log 2>&1; echo "kb rc=$?"; tail -6 gpurun_out/kb47.log | cut -c1-150
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_bwd_tma -s 2 -c 1 -o gpurun_out/scan_bwd_v1 -f python tools/kernel_bench.py --only bwd --batch 32 --dtype bf16 > gpurun_out/ncu_bwd.log 2>&1; echo "ncu rc=$?"
