#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
AUM_GEMM_LITE=1 AUM_SCAN_TMA_CH=160 timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_mixer_gpu.py tests/test_model_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/t_k.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|Error" gpurun_out/t_k.log | tail -8
AUM_GEMM_LITE=1 AUM_SCAN_TMA_CH=160 timeout 300 python tools/kernel_bench.py --only scan,gemm > gpurun_out/kb35.log 2>&1; echo "kb rc=$?"; grep -E "ch128|gemm_" gpurun_out/kb35.log | cut -c1-120
for cfg in "0 160" "1 160" "1 128"; do set -- $cfg
AUM_GEMM_LITE=$1 AUM_SCAN_TMA_CH=$2 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench35_$1_$2.json 2> gpurun_out/bench35.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench35_$1_$2.json')); print('lite=$1 ch=$2', {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'])"
done
