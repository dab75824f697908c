#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|Error" gpurun_out/t_all.log | tail -8
timeout 900 python bench.py > gpurun_out/bench44.json 2> gpurun_out/bench44.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench44.json')); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline'])"; tail -3 gpurun_out/bench44.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
