#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "scan" -p no:cacheprovider > gpurun_out/t_scan.log 2>&1; echo "scan rc=$?"
tail -15 gpurun_out/t_scan.log
timeout 900 python -m pytest tests/test_mixer_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/t_mixer.log 2>&1; echo "mixer rc=$?"
tail -5 gpurun_out/t_mixer.log
timeout 300 python tools/kernel_bench.py --only scan > gpurun_out/kb_scan.log 2>&1; echo "kb rc=$?"; cat gpurun_out/kb_scan.log
