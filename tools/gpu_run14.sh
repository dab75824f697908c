#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|Error" gpurun_out/t_all.log | tail -5
timeout 300 python tools/kernel_bench.py --only bwd --batch 32 --dtype bf16 > gpurun_out/kb_bwd.log 2>&1; echo "kb rc=$?"; cat gpurun_out/kb_bwd.log | tail -6
timeout 600 python tools/train_bench.py --steps 3 --warmup 2 --batch 32 > gpurun_out/train3.json 2> gpurun_out/train3.err; echo "train rc=$?"; cat gpurun_out/train3.json
