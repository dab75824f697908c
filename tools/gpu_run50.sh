#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -x -k "gemm" > gpurun_out/t_g.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_g.log | cut -c1-200
for e4 in 1 0; do for b in 64 32; do
AUM_GEMM_PAIR_EPI4=$e4 timeout 200 python tools/kernel_bench.py --only gemm --batch $b > gpurun_out/kb50_${e4}_$b.log 2>&1; echo "epi4=$e4 batch=$b rc=$?"; grep -E "in_proj\+silu" gpurun_out/kb50_${e4}_$b.log | cut -c1-110
done; done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench50.json 2> gpurun_out/bench50.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench50.json')); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'])"
AUM_GEMM_PAIR=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench50b.json 2> gpurun_out/bench50.err; echo "bench(nopair) rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench50b.json')); print({k:d[k] for k in ('value','ms_per_step')})"
