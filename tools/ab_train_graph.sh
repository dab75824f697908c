#!/bin/bash
# graph-captured training step + shadow weights + packed softplus epilogue: tests, then A/B
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_optim_gpu.py tests/test_backward_gpu.py tests/test_kernels_gpu.py tests/test_wgrad_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/t_graph.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|Error|FAILED" gpurun_out/t_graph.log | tail -15
for cfg in "0 0" "0 1" "1 1"; do
  set -- $cfg
  timeout 600 python tools/train_bench.py --steps 6 --warmup 3 --batch 32 --graph $1 --shadow $2 > gpurun_out/train_g$1_s$2.json 2> gpurun_out/train_g$1_s$2.err; echo "train graph=$1 shadow=$2 rc=$?"; cut -c1-330 gpurun_out/train_g$1_s$2.json; tail -3 gpurun_out/train_g$1_s$2.err
done
timeout 300 python tools/kernel_bench.py --only gemm > gpurun_out/kb_gemm.jsonl 2>&1; echo "kb rc=$?"; grep -E "dt_proj|in_proj" gpurun_out/kb_gemm.jsonl | cut -c1-200
