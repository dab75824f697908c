#!/bin/bash
# backward-scan specialisation: parity tests of everything that differentiates, then kernel + training-step A/B
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_backward_gpu.py tests/test_parity_tiers_gpu.py tests/test_optim_gpu.py tests/test_reference_dropin.py -q -m gpu -p no:cacheprovider > gpurun_out/t_bwd_spec.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|Error|FAILED" gpurun_out/t_bwd_spec.log | tail -15
for v in spec nospec; do
  if [ $v = nospec ]; then export AUM_SCAN_BWD_NOSPEC=1; else unset AUM_SCAN_BWD_NOSPEC; fi
  timeout 300 python tools/kernel_bench.py --only bwd --batch 32 --dtype bf16 > gpurun_out/kb_bwd_$v.jsonl 2>&1; echo "kb $v rc=$?"; cut -c1-200 gpurun_out/kb_bwd_$v.jsonl
  timeout 600 python tools/train_bench.py --steps 5 --warmup 3 --batch 32 > gpurun_out/train_$v.json 2> gpurun_out/train_$v.err; echo "train $v rc=$?"; cut -c1-260 gpurun_out/train_$v.json
done
