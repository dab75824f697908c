#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
for occ in 2 3; do AUM_SCAN_OCC=$occ timeout 300 python tools/kernel_bench.py --only scan > gpurun_out/kb_occ$occ.log 2>&1; echo "kb occ$occ rc=$?"; grep ch128 gpurun_out/kb_occ$occ.log; done
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_all.log | grep -vE "Warning|autocast|^$"
for occ in 2 3; do AUM_SCAN_OCC=$occ timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench5_occ$occ.json 2> gpurun_out/bench5.err; echo "bench occ$occ rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench5_occ$occ.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['avg_launch_ms'])"; done
AUM_SCAN_OCC=3 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench5_nograph.json 2>> gpurun_out/bench5.err; python -c "
import json; d=json.load(open('gpurun_out/bench5_nograph.json')); print('nograph', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])"
tail -5 gpurun_out/bench5.err
