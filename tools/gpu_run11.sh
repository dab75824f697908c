#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
for mb in 1 2 4; do timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --micro-batches $mb > gpurun_out/bench_mb$mb.json 2> gpurun_out/bench_mb.err; echo "bench mb$mb rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_mb$mb.json')); print('mb$mb', {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'])"; done
tail -3 gpurun_out/bench_mb.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --micro-batches 2 --no-graph > gpurun_out/bench_mb2_nograph.json 2>> gpurun_out/bench_mb.err; python -c "
import json; d=json.load(open('gpurun_out/bench_mb2_nograph.json')); print('mb2 nograph', {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 300 python -m pytest tests/test_model_gpu.py -q -m gpu -p no:cacheprovider -k "golden" 2>&1 | tail -2
