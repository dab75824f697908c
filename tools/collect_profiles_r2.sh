#!/bin/bash
# Round-2 evidence: bench line, ncu launch list of the same command, ncu --set full of the dominant kernel at the size
# bench.py launches it (32 sequences), per-kernel timings, config-5 sweep.  Outputs under gpurun_out/ (copied to profiles/).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2_bench_final.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_fwd.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/ncu_fwd.log 2>&1; echo "ncu launches rc=$?"; wc -l gpurun_out/r2_launches_fwd.csv
PREGATED=1 B=32 timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_fwd_tma -s 2 -c 1 -o gpurun_out/r2_scan_fwd_b32 -f python tools/scan_once.py > gpurun_out/ncu_scan.log 2>&1; echo "ncu scan rc=$?"
timeout 300 python tools/kernel_bench.py > gpurun_out/r2_kernel_bench.jsonl 2>&1; echo "kb rc=$?"
timeout 300 python tools/kernel_bench.py --only bwd --batch 32 --dtype bf16 >> gpurun_out/r2_kernel_bench.jsonl 2>&1
timeout 300 python tools/kernel_bench.py --only scan --batch 32 >> gpurun_out/r2_kernel_bench.jsonl 2>&1
timeout 600 python tools/scan_sweep.py > gpurun_out/r2_scan_sweep_config5.jsonl 2> gpurun_out/scan_sweep.err; echo "sweep rc=$?"; wc -l gpurun_out/r2_scan_sweep_config5.jsonl
