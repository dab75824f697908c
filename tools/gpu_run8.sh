#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 900 python -m pytest tests/test_backward_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/t_bwd.log 2>&1; echo "pytest bwd rc=$?"; grep -E "passed|failed|Error|error" gpurun_out/t_bwd.log | tail -30
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --deselect tests/test_backward_gpu.py > gpurun_out/t_all.log 2>&1; echo "pytest rest rc=$?"; tail -3 gpurun_out/t_all.log | grep -vE "Warning|autocast|^$"
