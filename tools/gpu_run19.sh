#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
AUM_SCAN_CH=128 timeout 300 python tools/kernel_bench.py --only gemm > gpurun_out/kb.log 2>&1; echo "kb rc=$?"; grep -E "proj" gpurun_out/kb.log
for pg in 0 1 0 1; do AUM_PREGATE_Z=$pg timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench10_pg$pg.json 2> gpurun_out/bench10.err; echo "bench pg$pg rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench10_pg$pg.json')); print('pregate=$pg', {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'])"; done
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_mixer_gpu.py tests/test_model_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -2
PREGATED=1 B=32 timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_fwd_tma -s 2 -c 1 -o gpurun_out/scan_v4 -f python tools/scan_once.py > gpurun_out/ncu_scan4.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --micro-batches 1 > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches rc=$?"; wc -l gpurun_out/launches_r1c.csv
