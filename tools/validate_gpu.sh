#!/bin/bash
# final validation of the committed tree: GPU tests, smoke, bench line
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|Error" gpurun_out/t_all.log | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_validate.json 2> gpurun_out/bench_validate.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_validate.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['cpu_baseline']['value'], d['clocks']); print(d['train']['value'], d['train']['ms_per_step']); print(d['extra'][0]['value'])"
