#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_backward_gpu.py tests/test_optim_gpu.py tests/test_parity_tiers_gpu.py tests/test_reference_dropin.py -q -m gpu -p no:cacheprovider > gpurun_out/t_grad16.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|Error|FAILED" gpurun_out/t_grad16.log | tail -15
timeout 300 python tools/kernel_bench.py --only bwd --batch 32 --dtype bf16 2>&1 | cut -c1-200
for v in 1 0; do
  AUM_GRAD_16BIT=$v timeout 600 python tools/train_bench.py --steps 6 --warmup 3 --batch 32 --graph 1 2>/dev/null | cut -c1-230 | sed "s/^/grad16=$v: /"
done
