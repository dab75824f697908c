#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 900 python -m pytest tests/test_backward_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/t_bwd.log 2>&1; echo "pytest bwd rc=$?"; grep -E "passed|failed" gpurun_out/t_bwd.log | tail -3
timeout 600 python tools/train_bench.py --steps 3 --warmup 2 --batch 32 > gpurun_out/train1.json 2> gpurun_out/train1.err; echo "train rc=$?"; cat gpurun_out/train1.json; tail -3 gpurun_out/train1.err
timeout 600 python tools/train_bench.py --steps 3 --warmup 2 --batch 32 --depth 2 > gpurun_out/train_d2.json 2>> gpurun_out/train1.err; cat gpurun_out/train_d2.json
