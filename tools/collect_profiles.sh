#!/bin/bash
# final round-1 evidence: bench line, launch list of the same command, ncu --set full of the dominant kernels, config-5 sweep
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; cat gpurun_out/bench_final.json | cut -c1-400
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2709 --launch-count 780 --csv --log-file gpurun_out/launches_fwd.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_fwd.log 2>&1; echo "ncu launches rc=$?"; wc -l gpurun_out/launches_fwd.csv
PREGATED=1 B=64 timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_fwd_tma -s 2 -c 1 -o gpurun_out/scan_fwd -f python tools/scan_once.py > gpurun_out/ncu_scan.log 2>&1; echo "ncu scan rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 83 -c 1 -o gpurun_out/gemm_dt_proj -f python tools/kernel_bench.py --only gemm --batch 64 > gpurun_out/ncu_dt_proj.log 2>&1; echo "ncu dt rc=$?"
timeout 600 python tools/scan_sweep.py > gpurun_out/scan_sweep.jsonl 2> gpurun_out/scan_sweep.err; echo "sweep rc=$?"; wc -l gpurun_out/scan_sweep.jsonl
timeout 300 python tools/kernel_bench.py > gpurun_out/kernel_bench.log 2>&1; echo "kb rc=$?"
timeout 600 python tools/train_bench.py --steps 3 --warmup 2 --batch 32 > gpurun_out/train_bench.json 2> gpurun_out/train_bench.err; echo "train rc=$?"; cat gpurun_out/train_bench.json | cut -c1-300
