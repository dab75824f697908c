#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for s in 0 1000 2500 4000 8000; do
AUM_SCAN_STAGGER=$s timeout 300 python tools/kernel_bench.py --only scan > gpurun_out/kb38_$s.log 2>&1; echo "stagger=$s rc=$?"; grep "pregated_ch128\|uniscan_ch128" gpurun_out/kb38_$s.log | cut -c1-110
done
