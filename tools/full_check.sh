#!/bin/bash
# whole GPU suite + the default bench line + a training-step op table
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|Error|FAILED" gpurun_out/t_all.log | tail -15
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_full.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['cpu_baseline']['value'], d['clocks']); print(d['train']['value'], d['train']['ms_per_step'], d['train'].get('launch_mode'), d['train']['gpu_launches']); print(d['extra'][0]['value'])"
tail -3 gpurun_out/bench_full.err
timeout 300 python tools/train_ops_profile.py > gpurun_out/r2_train_ops.txt 2>&1; echo "ops rc=$?"
