#!/bin/bash
# A/B of the 16-bit delta path: parity tests on the default (16-bit delta) build, then forward bench with and without it.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/t_delta16.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|Error|FAILED" gpurun_out/t_delta16.log | tail -15
for v in 1 0; do
  AUM_DELTA_16BIT=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/bench_delta16_$v.json 2> gpurun_out/bench_delta16_$v.err; echo "bench delta16=$v rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_delta16_$v.json')); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['avg_launch_ms'], d['roofline']['share_of_step'], d['clocks'])"
done
timeout 300 python tools/kernel_bench.py > gpurun_out/kb_delta16.jsonl 2>&1; echo "kb rc=$?"; tail -20 gpurun_out/kb_delta16.jsonl | cut -c1-250
