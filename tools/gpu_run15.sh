#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|Error" gpurun_out/t_all.log | tail -8
timeout 600 python tools/train_bench.py --steps 3 --warmup 2 --batch 32 > gpurun_out/train4.json 2> gpurun_out/train4.err; echo "train rc=$?"; cat gpurun_out/train4.json
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench6.json 2> gpurun_out/bench6.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench6.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'], d['clocks'])"
python __graft_entry__.py smoke 2>&1 | tail -2
