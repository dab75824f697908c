#!/bin/bash
# launch list of one training step (2-block AuM-Base, bf16) + torch op table + backward-scan kernel bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_train.csv python tools/train_bench.py --steps 1 --warmup 1 --batch 32 --depth 2 > gpurun_out/ncu_train.log 2>&1; echo "ncu train rc=$?"; wc -l gpurun_out/r2_launches_train.csv
timeout 300 python tools/train_ops_profile.py > gpurun_out/r2_train_ops.txt 2>&1; echo "ops rc=$?"
timeout 300 python tools/kernel_bench.py --only bwd --batch 32 --dtype bf16 > gpurun_out/kb_bwd.jsonl 2>&1; echo "kb rc=$?"; cut -c1-200 gpurun_out/kb_bwd.jsonl
