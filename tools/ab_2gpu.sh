#!/bin/bash
# 2 GPUs: the graph-captured training step with its NCCL all-reduce pieces inside the capture, vs eager
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for g in 1 0; do
  AUM_TRAIN_GRAPH=$g timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$g bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu_g$g.json 2> gpurun_out/bench_2gpu_g$g.err; echo "bench 2gpu graph=$g rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_2gpu_g$g.json')); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value']); t=d['train']; print(t['value'], t['ms_per_step'], t['launch_mode'][:30], t['allreduce'])"
  grep -iE "warn|error|fail" gpurun_out/bench_2gpu_g$g.err | head -5
done
