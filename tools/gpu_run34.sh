#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
AUM_SCAN_TMA_CH=192 timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_mixer_gpu.py -q -m gpu -p no:cacheprovider -x -k "scan or mixer or mamba" > gpurun_out/t_k.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|Error" gpurun_out/t_k.log | tail -8
for n in 3 4; do
AUM_SCAN_TMA_CH=192 AUM_SCAN_NSTG=$n timeout 300 python tools/kernel_bench.py --only scan > gpurun_out/kb34_$n.log 2>&1; echo "kb nstg=$n rc=$?"; grep ch128 gpurun_out/kb34_$n.log | cut -c1-120
AUM_SCAN_TMA_CH=192 AUM_SCAN_NSTG=$n timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench34_$n.json 2> gpurun_out/bench34.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench34_$n.json')); print('nstg=$n', {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'])"
done
