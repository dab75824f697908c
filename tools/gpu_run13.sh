#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 300 python tools/kernel_bench.py --only bwd --batch 32 --dtype bf16 > gpurun_out/kb_bwd.log 2>&1; echo "kb rc=$?"; cat gpurun_out/kb_bwd.log | tail -8
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_train.csv python tools/train_bench.py --steps 1 --warmup 0 --batch 32 --depth 2 > gpurun_out/ncu_train.log 2>&1; echo "ncu train rc=$?"; wc -l gpurun_out/launches_train.csv
