#!/bin/bash
# final validation of the committed tree: GPU tests, bench line, smoke, training step
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|Error" gpurun_out/t_all.log | tail -4
timeout 900 python bench.py > gpurun_out/bench_validate.json 2> gpurun_out/bench_validate.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_validate.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['share_of_step'], d['roofline']['xu_pipe']['frac'], d['cpu_baseline']['value'], d['clocks'])"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python tools/train_bench.py --steps 3 --warmup 2 --batch 32 > gpurun_out/train_validate.json 2> gpurun_out/train_validate.err; echo "train rc=$?"; cat gpurun_out/train_validate.json | cut -c1-260
