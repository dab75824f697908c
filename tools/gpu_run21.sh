#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|Error" gpurun_out/t_all.log | tail -8
AUM_SCAN_CH=128 timeout 300 python tools/kernel_bench.py --only scan > gpurun_out/kb.log 2>&1; echo "kb rc=$?"; grep -E "ch128" gpurun_out/kb.log
timeout 300 python tools/kernel_bench.py --only bwd --batch 32 --dtype bf16 > gpurun_out/kb_bwd.log 2>&1; grep -E "fwd_train" gpurun_out/kb_bwd.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench11.json 2> gpurun_out/bench11.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench11.json')); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'])"
