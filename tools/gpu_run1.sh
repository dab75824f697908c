#!/bin/bash
# First GPU bring-up: build, parity tests (tcgen05 isolated so a trap there cannot mask the rest), kernel timings.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "not tcgen05" -p no:cacheprovider > gpurun_out/t_kernels.log 2>&1; echo "kernels rc=$?"
tail -5 gpurun_out/t_kernels.log
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "tcgen05" -p no:cacheprovider > gpurun_out/t_tcgen05.log 2>&1; echo "tcgen05 rc=$?"
tail -5 gpurun_out/t_tcgen05.log
timeout 900 python -m pytest tests/test_mixer_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/t_mixer.log 2>&1; echo "mixer rc=$?"
tail -5 gpurun_out/t_mixer.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 300 python tools/kernel_bench.py --only conv,norm,scan > gpurun_out/kb_stream.log 2>&1; echo "kb1 rc=$?"; cat gpurun_out/kb_stream.log
timeout 300 python tools/kernel_bench.py --only gemm > gpurun_out/kb_gemm.log 2>&1; echo "kb2 rc=$?"; cat gpurun_out/kb_gemm.log
