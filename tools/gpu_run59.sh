#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/bench59_n4.json 2> gpurun_out/bench59_n4.err; echo "bench n4 rc=$?"; cat gpurun_out/bench59_n4.json | cut -c1-260; tail -2 gpurun_out/bench59_n4.err | cut -c1-200
