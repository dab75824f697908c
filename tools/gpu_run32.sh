#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|Error" gpurun_out/t_all.log | tail -8
timeout 300 python tools/kernel_bench.py --only conv,norm > gpurun_out/kb32.log 2>&1; echo "kb rc=$?"; tail -3 gpurun_out/kb32.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench32.json 2> gpurun_out/bench32.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench32.json')); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'])"
