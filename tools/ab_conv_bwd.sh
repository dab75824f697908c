#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_backward_gpu.py tests/test_optim_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/t_convbwd.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|Error|FAILED" gpurun_out/t_convbwd.log | tail -15
for v in 0 1 tile; do
  if [ $v = tile ]; then export AUM_CONV_BWD_TILE=1; else export AUM_CONV_BWD_VARIANT=$v; fi
  timeout 300 python tools/kernel_bench.py --only bwd --batch 32 --dtype bf16 2>&1 | grep conv1d | cut -c1-200 | sed "s/^/variant $v: /"
done
unset AUM_CONV_BWD_TILE AUM_CONV_BWD_VARIANT
timeout 600 python tools/train_bench.py --steps 6 --warmup 3 --batch 32 --graph 1 > gpurun_out/train_convstream.json 2> gpurun_out/train_convstream.err; echo "train rc=$?"; cut -c1-330 gpurun_out/train_convstream.json
AUM_CONV_BWD_VARIANT=1 timeout 600 python tools/train_bench.py --steps 6 --warmup 3 --batch 32 --graph 1 2>/dev/null | cut -c1-200
