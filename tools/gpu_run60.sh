#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2709 --launch-count 780 --csv --log-file gpurun_out/launches_fwd_v5.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_fwd5.log 2>&1; echo "ncu launches rc=$?"; wc -l gpurun_out/launches_fwd_v5.csv
timeout 300 python tools/kernel_bench.py > gpurun_out/kb60.log 2>&1; echo "kb rc=$?"
