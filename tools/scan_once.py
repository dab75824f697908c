"""Launch the fused bidirectional scan a few times at config-2 shape (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "audio-mamba-aum_b200"))
import torch
from aum_b200 import ops
dev = "cuda"; B, Lq, Di, N = int(os.environ.get("B", 64)), 513, 1536, 16
dt = torch.float16
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *sh, dtype=dt: torch.randn(*sh, device=dev, generator=g).to(dtype)
u, z = rn(B, Lq, Di), rn(B, Lq, Di)
delta = torch.nn.functional.softplus(rn(B, Lq, Di, dtype=torch.float32) - 2.0)
if os.environ.get("DELTA16", "1") == "1":      # as the inference path stores it (mixer._DELTA_16BIT)
    delta = delta.to(dt)
bc = rn(B, Lq, 2 * N, dtype=torch.float32)
A = -torch.exp(torch.log(torch.arange(1, N + 1, device=dev, dtype=torch.float32)).repeat(Di, 1) + 0.1 * rn(Di, N, dtype=torch.float32))
A_b = -torch.exp(torch.log(torch.arange(1, N + 1, device=dev, dtype=torch.float32)).repeat(Di, 1) + 0.1 * rn(Di, N, dtype=torch.float32))
Dv = torch.ones(Di, device=dev)
mk = lambda Ax: ops.ScanDirection(u, delta, Ax, bc[..., :N], bc[..., N:], Dv)
out = torch.empty_like(u)
for _ in range(4):
    ops.selective_scan(mk(A), mk(A_b), z, out=out, z_pregated=os.environ.get("PREGATED", "0") == "1")
torch.cuda.synchronize()
print("ok")
