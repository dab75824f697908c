#!/bin/bash
# Round-2 final evidence (outputs under gpurun_out/, summaries copied to profiles/ afterwards):
#   whole GPU suite, smoke(), the default bench line, ncu launch lists of the forward and of a training step, ncu --set full of
#   the forward scan at the size bench.py launches it (-> profiles/r2_scan_traffic.json) and of the backward scan as the
#   training step calls it, per-kernel timings, config-5 sweep.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/t_final.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|Error|FAILED" gpurun_out/t_final.log | tail -8
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
PREGATED=1 DELTA16=1 B=32 timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_fwd_tma -s 2 -c 1 -o gpurun_out/r2_scan_fwd_b32 -f python tools/scan_once.py > gpurun_out/ncu_scan.log 2>&1; echo "ncu scan rc=$?"
B=32 DT=bf16 timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_bwd_tma -s 2 -c 1 -o gpurun_out/r2_scan_bwd_v3 -f python tools/one_kernel.py scan_bwd > gpurun_out/ncu_bwd.log 2>&1; echo "ncu bwd rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_fwd.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/ncu_fwd.log 2>&1; echo "ncu launches fwd rc=$?"; wc -l gpurun_out/r2_launches_fwd.csv
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_train.csv python tools/train_bench.py --steps 1 --warmup 1 --batch 32 --depth 2 > gpurun_out/ncu_train.log 2>&1; echo "ncu launches train rc=$?"; wc -l gpurun_out/r2_launches_train.csv
timeout 300 python tools/kernel_bench.py > gpurun_out/r2_kernel_bench.jsonl 2>&1; echo "kb rc=$?"
timeout 300 python tools/kernel_bench.py --only bwd --batch 32 --dtype bf16 >> gpurun_out/r2_kernel_bench.jsonl 2>&1
timeout 300 python tools/kernel_bench.py --only scan --batch 32 >> gpurun_out/r2_kernel_bench.jsonl 2>&1
timeout 600 python tools/scan_sweep.py > gpurun_out/r2_scan_sweep_config5.jsonl 2> gpurun_out/scan_sweep.err; echo "sweep rc=$?"; wc -l gpurun_out/r2_scan_sweep_config5.jsonl
timeout 300 python tools/train_ops_profile.py > gpurun_out/r2_train_ops.txt 2>&1; echo "ops rc=$?"
