#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for mb in 1 2 4; do
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --micro-batches $mb > gpurun_out/bench37_$mb.json 2> gpurun_out/bench37.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench37_$mb.json')); print('mb=$mb', {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'])"
done
