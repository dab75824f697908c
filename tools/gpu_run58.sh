#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for pa in 0 1; do for b in 64 32; do
AUM_GEMM_PAIR_ACT=$pa timeout 200 python tools/kernel_bench.py --only gemm --batch $b > gpurun_out/kb58_${pa}_$b.log 2>&1; echo "pair_act=$pa batch=$b rc=$?"; grep -E "in_proj\+silu" gpurun_out/kb58_${pa}_$b.log | cut -c1-105
done; done
for pa in 0 1; do
AUM_GEMM_PAIR_ACT=$pa timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench58_$pa.json 2> gpurun_out/bench58.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench58_$pa.json')); print('pair_act=$pa', {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'])"
done
