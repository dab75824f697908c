#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|Error" gpurun_out/t_all.log | tail -8
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench27.json 2> gpurun_out/bench27.err; echo "bench rc=$?"; cat gpurun_out/bench27.json; tail -3 gpurun_out/bench27.err
timeout 300 python tools/kernel_bench.py > gpurun_out/kb27.log 2>&1; echo "kb rc=$?"; tail -40 gpurun_out/kb27.log
