#!/bin/bash
# bench.py on N GPUs of one box exactly as the driver launches it:  bash tools/bench_ngpu.sh N
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=${1:-2}
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err; echo "bench ${N}gpu rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_${N}gpu.json')); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d['clocks']); t=d['train']; print(t['value'], t['ms_per_step'], t['launch_mode'][:30], t['allreduce']); print(d['extra'][0]['value'])"
grep -iE "warn|error|fail" gpurun_out/r2_bench_${N}gpu.err | head -5
