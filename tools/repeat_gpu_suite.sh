#!/bin/bash
# flakiness check: the whole GPU suite three times in a row (fresh processes), then the bench twice
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for i in 1 2 3; do
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/t_rep$i.log 2>&1; echo "pytest run $i rc=$?"; grep -E "passed|failed|Error" gpurun_out/t_rep$i.log | tail -2
done
for i in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_repeat_$i.json 2> gpurun_out/bench_repeat.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_repeat_$i.json')); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['clocks'])"
done
