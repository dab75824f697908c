#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_backward_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/t_bwd.log 2>&1; echo "pytest bwd rc=$?"; tail -5 gpurun_out/t_bwd.log | cut -c1-300
timeout 300 python tools/kernel_bench.py --only bwd --batch 32 --dtype bf16 > gpurun_out/kb51.log 2>&1; echo "kb rc=$?"; tail -5 gpurun_out/kb51.log | cut -c1-150
timeout 600 python tools/train_bench.py --steps 3 --warmup 2 --batch 32 > gpurun_out/train51.json 2> gpurun_out/train51.err; echo "train rc=$?"; cat gpurun_out/train51.json | cut -c1-300; tail -2 gpurun_out/train51.err
