#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_optim_gpu.py tests/test_backward_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/t_opt.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/t_opt.log | cut -c1-300
timeout 600 python tools/train_bench.py --steps 3 --warmup 2 --batch 32 > gpurun_out/train53.json 2> gpurun_out/train53.err; echo "train rc=$?"; cat gpurun_out/train53.json | cut -c1-300; tail -2 gpurun_out/train53.err
timeout 600 python tools/train_bench.py --steps 3 --warmup 2 --batch 32 --torch-adam > gpurun_out/train53t.json 2> gpurun_out/train53t.err; echo "train(torch adam) rc=$?"; cat gpurun_out/train53t.json | cut -c1-300
