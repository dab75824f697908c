"""BASELINE config 5: selective-scan sweep L in {128, 512, 1024, 2048, 4096}, d_state 16, batch 32, D = 1536 —
fused forward+reverse scan and single-direction scan, fp16 and fp32 I/O; achieved algorithmic GB/s vs the measured
HBM peak, plus T exp/s vs the measured MUFU ceiling.  CUDA events, L2 flushed between iterations.
    python tools/scan_sweep.py > profiles/r1_scan_sweep.jsonl"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "audio-mamba-aum_b200"))
import torch  # noqa: E402
from aum_b200 import ops  # noqa: E402

try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0
dev = "cuda"
B, D, N = 32, 1536, 16
flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=dev)
g = torch.Generator(device=dev).manual_seed(3949)


def timeit(fn, iters=8):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return sorted(ts)[len(ts) // 2]


for dt in (torch.float16, torch.float32):
    s = 2 if dt == torch.float16 else 4
    for Lq in (128, 512, 1024, 2048, 4096):
        rn = lambda *sh, dtype=dt: torch.randn(*sh, device=dev, generator=g).to(dtype)
        u, z = rn(B, Lq, D), rn(B, Lq, D)
        delta = torch.nn.functional.softplus(rn(B, Lq, D, dtype=torch.float32) - 2.0)
        bc = rn(B, Lq, 2 * N, dtype=torch.float32)
        mkA = lambda: -torch.exp(torch.log(torch.arange(1, N + 1, device=dev, dtype=torch.float32)).repeat(D, 1)
                                 + 0.1 * rn(D, N, dtype=torch.float32))
        A, A_b = mkA(), mkA()
        Dv = torch.ones(D, device=dev)
        out = torch.empty_like(u)
        mk = lambda Ax: ops.ScanDirection(u, delta, Ax, bc[..., :N], bc[..., N:], Dv)
        alg = B * Lq * D * (3 * s + 4) + B * Lq * 2 * N * 4
        for name, f, nd in (("bidirectional", lambda: ops.selective_scan(mk(A), mk(A_b), z, out=out), 2),
                            ("forward-only", lambda: ops.selective_scan(mk(A), None, z, out=out), 1)):
            ms = timeit(f)
            print(json.dumps({"scan": name, "dtype": str(dt).split(".")[-1], "B": B, "D": D, "N": N, "L": Lq, "ms": round(ms, 4),
                              "algorithmic_GBs": round(alg / ms / 1e6, 1), "hbm_frac": round(alg / ms / 1e6 / PEAK, 4),
                              "T_exp_per_s": round(B * Lq * D * 16 * nd / ms / 1e9, 3)}), flush=True)
