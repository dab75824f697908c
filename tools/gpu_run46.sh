#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench46_n2.json 2> gpurun_out/bench46_n2.err; echo "bench n2 rc=$?"; cat gpurun_out/bench46_n2.json | cut -c1-700; tail -2 gpurun_out/bench46_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/train_bench.py --steps 3 --warmup 2 --batch 32 > gpurun_out/train46_n2.json 2> gpurun_out/train46_n2.err; echo "train n2 rc=$?"; cat gpurun_out/train46_n2.json | cut -c1-400
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/ref46_n2.json 2> gpurun_out/ref46_n2.err; echo "ref n2 rc=$?"; cat gpurun_out/ref46_n2.json | cut -c1-300
