#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -x -k scan > gpurun_out/t_k.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|Error" gpurun_out/t_k.log | tail -8
timeout 300 python tools/kernel_bench.py --only scan > gpurun_out/kb30.log 2>&1; echo "kb rc=$?"; grep ch128 gpurun_out/kb30.log
