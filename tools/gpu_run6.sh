#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -x -k "scan" > gpurun_out/t_scan.log 2>&1; echo "pytest scan rc=$?"; tail -12 gpurun_out/t_scan.log | grep -vE "Warning|autocast|^$"
timeout 300 python tools/kernel_bench.py --only scan > gpurun_out/kb.log 2>&1; echo "kb rc=$?"; cat gpurun_out/kb.log
AUM_SCAN_GENERIC=1 timeout 300 python tools/kernel_bench.py --only scan > gpurun_out/kb_generic.log 2>&1; echo "kb generic rc=$?"; cat gpurun_out/kb_generic.log
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_all.log | grep -vE "Warning|autocast|^$"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench4.json 2> gpurun_out/bench4.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench4.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['avg_launch_ms'])"
