#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|Error" gpurun_out/t_all.log | tail -8
timeout 300 python tools/kernel_bench.py --only scan,gemm > gpurun_out/kb.log 2>&1; echo "kb rc=$?"; grep -E "ch128|in_proj" gpurun_out/kb.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench7.json 2> gpurun_out/bench7.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench7.json')); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'])"
timeout 300 python tools/scan_sweep.py > gpurun_out/scan_sweep.jsonl 2> gpurun_out/scan_sweep.err; echo "sweep rc=$?"; cat gpurun_out/scan_sweep.jsonl | cut -c1-200
