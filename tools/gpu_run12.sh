#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|Error" gpurun_out/t_all.log | tail -5
timeout 600 python tools/train_bench.py --steps 3 --warmup 2 --batch 32 > gpurun_out/train2.json 2> gpurun_out/train2.err; echo "train rc=$?"; cat gpurun_out/train2.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 260 --csv --log-file gpurun_out/launches_train.csv python tools/train_bench.py --steps 1 --warmup 1 --batch 32 --depth 2 > gpurun_out/ncu_train.log 2>&1; echo "ncu train rc=$?"
