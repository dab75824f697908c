#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
AUM_SCAN_NSTG=4 PREGATED=1 B=32 timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_fwd_tma -s 2 -c 1 -o gpurun_out/scan_v5 -f python tools/scan_once.py > gpurun_out/ncu_scan5.log 2>&1; echo "ncu rc=$?"
