#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench45.json 2> gpurun_out/bench45.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench45.json')); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline'])"; tail -3 gpurun_out/bench45.err
