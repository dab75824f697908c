#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python tools/kernel_bench.py --only bwd --batch 32 --dtype bf16 > gpurun_out/kb52.log 2>&1; echo "kb rc=$?"; grep biscan_bwd gpurun_out/kb52.log | cut -c1-150
