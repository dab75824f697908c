#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q -m gpu -p no:cacheprovider -x -k "patchify or audio_mamba or aum_ or cuda_graph" > gpurun_out/t_fe.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/t_fe.log | cut -c1-250
