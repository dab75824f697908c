#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/t_model.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/t_model.log | cut -c1-250
