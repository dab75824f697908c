#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|Error" gpurun_out/t_all.log | tail -8
AUM_SCAN_CH=128 timeout 300 python tools/kernel_bench.py --only scan,gemm > gpurun_out/kb.log 2>&1; echo "kb rc=$?"; grep -E "ch128|in_proj" gpurun_out/kb.log
for pg in 0 1; do AUM_PREGATE_Z=$pg timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench8_pg$pg.json 2> gpurun_out/bench8.err; echo "bench pg$pg rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench8_pg$pg.json')); print('pregate=$pg', {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'])"; done
