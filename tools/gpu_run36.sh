#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|Error" gpurun_out/t_all.log | tail -8
timeout 600 python tools/train_bench.py --steps 3 --warmup 2 --batch 32 > gpurun_out/train36.json 2> gpurun_out/train36.err; echo "train rc=$?"; cat gpurun_out/train36.json; tail -3 gpurun_out/train36.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_train.csv python tools/train_bench.py --steps 1 --warmup 1 --batch 32 --depth 2 > gpurun_out/ncu_train.log 2>&1; echo "ncu train rc=$?"; wc -l gpurun_out/launches_train.csv
