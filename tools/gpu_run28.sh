#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PREGATED=1 B=64 timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_fwd_tma -s 2 -c 1 -o gpurun_out/scan_v9 -f python tools/scan_once.py > gpurun_out/ncu_scan9.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_scan6.log
