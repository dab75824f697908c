#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_n2.json')); print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}, d['e2e']['value'], d['clocks'])"; tail -3 gpurun_out/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_ref_n2.json 2>> gpurun_out/bench_n2.err; echo "ref n2 rc=$?"; cut -c1-300 gpurun_out/bench_ref_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/train_bench.py --steps 3 --warmup 2 --batch 32 > gpurun_out/train_n2.json 2> gpurun_out/train_n2.err; echo "train n2 rc=$?"; cat gpurun_out/train_n2.json; tail -3 gpurun_out/train_n2.err
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench n1 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d['cpu_baseline'], d['roofline']['frac'], d['roofline']['traffic'])"
