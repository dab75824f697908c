"""Launch ONE kernel of the path a few times at its config-2 / config-3 shape (for `ncu -k regex:... -s 2 -c 1`).
    python tools/one_kernel.py dt_proj16 | dt_proj32 | conv1d_bwd | conv_xproj | scan_bwd"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "audio-mamba-aum_b200"))
import torch
from aum_b200 import _lib as L, ops
which = sys.argv[1]
dev = "cuda"
B, Lq, Dm, Di, N, R = int(os.environ.get("B", 32)), 513, 768, 1536, 16, 48
M = B * Lq
dt = torch.bfloat16 if os.environ.get("DT", "fp16") == "bf16" else torch.float16
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *sh, dtype=dt, sc=1.0: (torch.randn(*sh, device=dev, generator=g) * sc).to(dtype)
f32 = dict(device=dev, dtype=torch.float32)
if which in ("dt_proj16", "dt_proj32"):
    a, w, bias = rn(M, R), rn(Di, R, sc=R ** -0.5), rn(Di, dtype=torch.float32)
    out = torch.empty(M, Di, device=dev, dtype=dt if which == "dt_proj16" else torch.float32)
    fn = lambda: ops.gemm_tn(a, w, out=out, k=R, bias=bias, act=L.ACT_SOFTPLUS, backend=L.GEMM_TCGEN05)
elif which == "conv1d_bwd":
    x, w, b_ = rn(B, Lq, 2 * Di), rn(Di, 4, dtype=torch.float32), rn(Di, dtype=torch.float32)
    g1, g2, g3 = (rn(B, Lq, Di, dtype=torch.float32) for _ in range(3))
    dx = torch.empty((B, Lq, Di), device=dev, dtype=dt)
    dw, db_ = torch.zeros((Di, 4), **f32), torch.zeros(Di, **f32)
    fn = lambda: ops.causal_conv1d_bwd(x[..., :Di], w, b_, g1, dx, dw, db_, dout2=g2, dout3=g3)
elif which == "conv_xproj":
    x, cw, cb = rn(B, Lq, 2 * Di), rn(Di, 4, dtype=torch.float32), rn(Di, dtype=torch.float32)
    wx = rn(R + 2 * N, Di, sc=Di ** -0.5)
    dtb, bc = torch.empty(M, 48, device=dev, dtype=dt), torch.empty(M, 2 * N, **f32)
    fn = lambda: ops.conv_xproj(x[..., :Di], cw, cb, wx, R, dtb, bc)
elif which == "scan_bwd":       # as the training step calls it: 16-bit delta / du / ddelta, per-direction outputs, softplus'
    u, z, ypre, dout = (rn(B, Lq, Di) for _ in range(4))
    delta = torch.nn.functional.softplus(rn(B, Lq, Di, dtype=torch.float32) - 2.0).to(dt)
    bc = rn(B, Lq, 2 * N, dtype=torch.float32)
    mkA = lambda: -torch.exp(torch.log(torch.arange(1, N + 1, device=dev, dtype=torch.float32)).repeat(Di, 1) + 0.1 * rn(Di, N, dtype=torch.float32))
    A, A_b = mkA(), mkA()
    Dv = torch.ones(Di, device=dev)
    h = dict(device=dev, dtype=dt)
    du, dd, du2, dd2 = (torch.empty((B, Lq, Di), **h) for _ in range(4))
    dbc = torch.zeros((B, Lq, 2 * N), **f32)
    dA, dAb, dD = torch.zeros((Di, N), **f32), torch.zeros((Di, N), **f32), torch.zeros(Di, **f32)
    dz, oz, out = torch.empty_like(u), torch.empty_like(u), torch.empty_like(u)
    ckf, ckb = ops.scan_bwd_workspace(B, Lq, Di, dev), ops.scan_bwd_workspace(B, Lq, Di, dev)
    ops.selective_scan(ops.ScanDirection(u, delta, A, bc[..., :N], bc[..., N:], Dv, ckpt=ckf),
                       ops.ScanDirection(u, delta, A_b, bc[..., :N], bc[..., N:], Dv, ckpt=ckb), z, out=out, y_pre=ypre)
    fn = lambda: ops.selective_scan_bwd(
        ops.ScanBwdDirection(u, delta, A, bc, Dv, du, dd, dA, dD, dbc, ckf, ckpt_valid=True),
        ops.ScanBwdDirection(u, delta, A_b, bc, Dv, du2, dd2, dAb, dD, dbc, ckb, ckpt_valid=True),
        z, ypre, dout, dz, oz, softplus_grad=True)
else:
    raise SystemExit("unknown kernel " + which)
for _ in range(4):
    fn()
torch.cuda.synchronize()
print("ok")
