#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
./tools/microbench > gpurun_out/microbench.txt 2>&1; echo "microbench rc=$?"; cat gpurun_out/microbench.txt
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/t_all.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench rc=$?"; cat gpurun_out/bench1.json; tail -3 gpurun_out/bench1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_fwd_kernel -s 2 -c 1 -f -o gpurun_out/scan_v2 python tools/scan_once.py > gpurun_out/ncu_scan.log 2>&1; echo "ncu scan rc=$?"; tail -3 gpurun_out/ncu_scan.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches rc=$?"; tail -2 gpurun_out/ncu_bench.log
