#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fwd_graph.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_fwd.log 2>&1; echo "ncu fwd rc=$?"; wc -l gpurun_out/launches_fwd_graph.csv; tail -2 gpurun_out/ncu_fwd.log | cut -c1-300
