#!/bin/bash
# the fused scan at launch sizes that fill whole waves (296 CTA slots = 2 x 148 SMs; 12 CTAs per sequence at Di = 1536)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for b in 74 148 32 64; do timeout 300 python tools/kernel_bench.py --only scan --batch $b 2>&1 | grep "delta16" | cut -c1-200 | sed "s/^/B=$b ($((b*12)) CTAs): /"; done | tee gpurun_out/r2_scan_waves.txt
