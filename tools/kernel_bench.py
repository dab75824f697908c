"""Per-kernel timing at AuM-Base shapes (config 2: B=64, L=513, Dm=768, Di=1536, R=48, N=16).
Usage: python tools/kernel_bench.py [--batch 64] [--dtype fp16] [--only scan,gemm,...]
CUDA-event timing, L2 flushed between iterations.  Prints one JSON line per kernel."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "audio-mamba-aum_b200"))
import torch  # noqa: E402

from aum_b200 import ops, _lib as L  # noqa: E402

PEAK_GBS, PEAK_TF = 6575.8, 1693.1
try:
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    PEAK_GBS, PEAK_TF = pk["hbm_gbs"], pk["bf16_tflops"]
except Exception:
    pass


def timeit(fn, iters=10, warmup=3, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--dtype", default="fp16")
    ap.add_argument("--only", default="")
    ap.add_argument("--L", type=int, default=513)
    args = ap.parse_args()
    dt = {"fp16": torch.float16, "bf16": torch.bfloat16, "fp32": torch.float32}[args.dtype]
    s = 4 if dt == torch.float32 else 2
    only = set(filter(None, args.only.split(",")))
    dev = "cuda"
    B, Lq, Dm, Di, R, N = args.batch, args.L, 768, 1536, 48, 16
    M = B * Lq
    flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2
    g = torch.Generator(device=dev).manual_seed(0)
    rn = lambda *sh, dtype=dt, sc=1.0: (sc * torch.randn(*sh, device=dev, generator=g)).to(dtype)

    def report(name, ms, bytes_=None, flops=None, **kw):
        r = {"kernel": name, "ms": round(ms, 4)}
        if bytes_:
            r["GBs"] = round(bytes_ / ms / 1e6, 1)
            r["hbm_frac"] = round(bytes_ / ms / 1e6 / PEAK_GBS, 3)
        if flops:
            r["TFs"] = round(flops / ms / 1e9, 1)
            r["tensor_frac"] = round(flops / ms / 1e9 / PEAK_TF, 3)
        r.update(kw)
        print(json.dumps(r), flush=True)

    if not only or "conv" in only:
        xz = rn(B, Lq, 2 * Di)
        w, b = rn(Di, 4, dtype=torch.float32), rn(Di, dtype=torch.float32)
        ms = timeit(lambda: ops.causal_conv1d(xz[..., :Di], w, b), flush=flush)
        report("conv1d_silu", ms, bytes_=2 * M * Di * s)
    if dt != torch.float32 and (not only or "fused" in only):
        # conv + SiLU fused into x_proj's operand producer (aum_conv_xproj_fwd): reads x, writes u, dt, B|C
        xz = rn(B, Lq, 2 * Di)
        w, b = rn(Di, 4, dtype=torch.float32), rn(Di, dtype=torch.float32)
        wx = rn(R + 2 * N, Di, sc=Di ** -0.5)
        dt_o = torch.empty(M, (R + 7) // 8 * 8, device=dev, dtype=dt)
        bc_o = torch.empty(M, 2 * N, device=dev, dtype=torch.float32)
        u_o = torch.empty(B, Lq, Di, device=dev, dtype=dt)
        ms = timeit(lambda: ops.conv_xproj(xz[..., :Di], w, b, wx, R, dt_o, bc_o, u=u_o), flush=flush)
        report("conv_xproj_fused", ms, bytes_=2 * M * Di * s + M * (R * s + 2 * N * 4))
    if not only or "norm" in only:
        x, r_ = rn(M, Dm), rn(M, Dm, dtype=torch.float32)
        w = torch.ones(Dm, device=dev)
        ms = timeit(lambda: ops.add_rmsnorm(x, w, None, r_, prenorm=True), flush=flush)
        report("add_rmsnorm", ms, bytes_=M * Dm * (2 * s + 8))
    if not only or "scan" in only:
        u, z = rn(B, Lq, Di), rn(B, Lq, Di)
        delta = torch.nn.functional.softplus(rn(B, Lq, Di, dtype=torch.float32) - 2.0)
        bc = rn(B, Lq, 2 * N, dtype=torch.float32)
        A = -torch.exp(torch.log(torch.arange(1, N + 1, device=dev, dtype=torch.float32)).repeat(Di, 1)
                       + 0.1 * rn(Di, N, dtype=torch.float32))
        A_b = -torch.exp(torch.log(torch.arange(1, N + 1, device=dev, dtype=torch.float32)).repeat(Di, 1)
                         + 0.1 * rn(Di, N, dtype=torch.float32))
        Dv = torch.ones(Di, device=dev)
        mk = lambda Ax: ops.ScanDirection(u, delta, Ax, bc[..., :N], bc[..., N:], Dv)
        out = torch.empty_like(u)
        alg = M * Di * (3 * s + 4) + M * 2 * N * 4     # u, z, out (s) + delta fp32 + B,C fp32
        for ch in ("64", "128"):
            os.environ["AUM_SCAN_CH"] = ch
            ms = timeit(lambda: ops.selective_scan(mk(A), mk(A_b), z, out=out), flush=flush)
            report(f"biscan_ch{ch}", ms, bytes_=alg, exp_per_s_T=round(M * Di * 32 / ms / 1e9, 3))
            ms = timeit(lambda: ops.selective_scan(mk(A), mk(A_b), z, out=out, z_pregated=True), flush=flush)
            report(f"biscan_pregated_ch{ch}", ms, bytes_=alg, exp_per_s_T=round(M * Di * 32 / ms / 1e9, 3))
            ms = timeit(lambda: ops.selective_scan(mk(A), None, z, out=out), flush=flush)
            report(f"uniscan_ch{ch}", ms, bytes_=alg, exp_per_s_T=round(M * Di * 16 / ms / 1e9, 3))
        if dt != torch.float32:       # delta in the activation dtype (the inference default): SURVEY 8(d)'s byte count exactly
            d16 = delta.to(dt)
            mk16 = lambda Ax: ops.ScanDirection(u, d16, Ax, bc[..., :N], bc[..., N:], Dv)
            ms = timeit(lambda: ops.selective_scan(mk16(A), mk16(A_b), z, out=out, z_pregated=True), flush=flush)
            report("biscan_pregated_delta16", ms, bytes_=M * Di * 4 * s + M * 2 * N * 4, exp_per_s_T=round(M * Di * 32 / ms / 1e9, 3))
    if "bwd" in only:
        u, z = rn(B, Lq, Di), rn(B, Lq, Di)
        ypre, dout = rn(B, Lq, Di), rn(B, Lq, Di)
        delta = torch.nn.functional.softplus(rn(B, Lq, Di, dtype=torch.float32) - 2.0)
        bc = rn(B, Lq, 2 * N, dtype=torch.float32)
        mkA = lambda: -torch.exp(torch.log(torch.arange(1, N + 1, device=dev, dtype=torch.float32)).repeat(Di, 1)
                                 + 0.1 * rn(Di, N, dtype=torch.float32))
        A, A_b = mkA(), mkA()
        Dv = torch.ones(Di, device=dev)
        f32 = dict(device=dev, dtype=torch.float32)
        du, dd = torch.empty((B, Lq, Di), **f32), torch.empty((B, Lq, Di), **f32)
        dbc = torch.zeros((B, Lq, 2 * N), **f32)
        dA, dAb, dD = torch.zeros((Di, N), **f32), torch.zeros((Di, N), **f32), torch.zeros(Di, **f32)
        dz, oz = torch.empty_like(u), torch.empty_like(u)
        ckf, ckb = ops.scan_bwd_workspace(B, Lq, Di, dev), ops.scan_bwd_workspace(B, Lq, Di, dev)
        out = torch.empty_like(u)
        # forward with checkpoints (what training does), then the backward with and without them
        fw = lambda: ops.selective_scan(ops.ScanDirection(u, delta, A, bc[..., :N], bc[..., N:], Dv, ckpt=ckf),
                                        ops.ScanDirection(u, delta, A_b, bc[..., :N], bc[..., N:], Dv, ckpt=ckb), z, out=out, y_pre=ypre)
        report("biscan_fwd_train(ckpt+ypre)", timeit(fw, flush=flush), exp_per_s_T=None)
        for valid in (True, False):
            bw = lambda: ops.selective_scan_bwd(
                ops.ScanBwdDirection(u, delta, A, bc, Dv, du, dd, dA, dD, dbc, ckf, ckpt_valid=valid),
                ops.ScanBwdDirection(u, delta, A_b, bc, Dv, du, dd, dAb, dD, dbc, ckb, ckpt_valid=valid),
                z, ypre, dout, dz, oz)
            report(f"biscan_bwd(ckpt_valid={valid})", timeit(bw, iters=5, flush=flush))
        # as the training step calls it (autograd.py): one du / ddelta pair per direction, softplus' folded in ->
        # the compile-time specialised instantiation of scan_bwd_tma_kernel
        du2, dd2 = torch.empty((B, Lq, Di), **f32), torch.empty((B, Lq, Di), **f32)
        bw = lambda: ops.selective_scan_bwd(
            ops.ScanBwdDirection(u, delta, A, bc, Dv, du, dd, dA, dD, dbc, ckf, ckpt_valid=True),
            ops.ScanBwdDirection(u, delta, A_b, bc, Dv, du2, dd2, dAb, dD, dbc, ckb, ckpt_valid=True),
            z, ypre, dout, dz, oz, softplus_grad=True)
        report("biscan_bwd(training call: per-direction du/ddelta, softplus')", timeit(bw, iters=5, flush=flush))
        if dt != torch.float32:
            h = dict(device=dev, dtype=dt)
            duh, ddh, du2h, dd2h = (torch.empty((B, Lq, Di), **h) for _ in range(4))
            bwh = lambda: ops.selective_scan_bwd(
                ops.ScanBwdDirection(u, delta, A, bc, Dv, duh, ddh, dA, dD, dbc, ckf, ckpt_valid=True),
                ops.ScanBwdDirection(u, delta, A_b, bc, Dv, du2h, dd2h, dAb, dD, dbc, ckb, ckpt_valid=True),
                z, ypre, dout, dz, oz, softplus_grad=True)
            report("biscan_bwd(training call, 16-bit du/ddelta)", timeit(bwh, iters=5, flush=flush))
            d16 = delta.to(dt)
            fw16 = lambda: ops.selective_scan(ops.ScanDirection(u, d16, A, bc[..., :N], bc[..., N:], Dv, ckpt=ckf),
                                              ops.ScanDirection(u, d16, A_b, bc[..., :N], bc[..., N:], Dv, ckpt=ckb), z, out=out, y_pre=ypre)
            report("biscan_fwd_train(ckpt+ypre, 16-bit delta)", timeit(fw16, flush=flush))
            bwh16 = lambda: ops.selective_scan_bwd(
                ops.ScanBwdDirection(u, d16, A, bc, Dv, duh, ddh, dA, dD, dbc, ckf, ckpt_valid=True),
                ops.ScanBwdDirection(u, d16, A_b, bc, Dv, du2h, dd2h, dAb, dD, dbc, ckb, ckpt_valid=True),
                z, ypre, dout, dz, oz, softplus_grad=True)
            report("biscan_bwd(training call, 16-bit du/ddelta/delta: %s CTAs per SM)" % ("3" if not os.environ.get("AUM_SCAN_BWD_4CTA") else "4"),
                   timeit(bwh16, iters=5, flush=flush))
            csum = torch.zeros(Di, **f32)
            report("sum_cast_colsum(fp32 in)", timeit(lambda: ops.sum_cast_colsum(dd.view(M, Di), dd2.view(M, Di), dt, csum), flush=flush),
                   bytes_=M * Di * (8 + s))
            report("sum_cast_colsum(16-bit in)", timeit(lambda: ops.sum_cast_colsum(ddh.view(M, Di), dd2h.view(M, Di), dt, csum), flush=flush),
                   bytes_=M * Di * (4 + s))
        dy_n, dro_n, r_n = rn(M, Dm), rn(M, Dm, dtype=torch.float32), rn(M, Dm, dtype=torch.float32)
        rstd_n, w_n, dwt_n = torch.rand(M, device=dev) + 0.5, torch.ones(Dm, device=dev), torch.zeros(Dm, device=dev)
        report("add_rmsnorm_bwd", timeit(lambda: ops.add_rmsnorm_bwd(dy_n, dro_n, r_n, rstd_n, w_n, dwt_n, want_dres_in=True), flush=flush),
               bytes_=M * Dm * (2 * s + 12))
        x = rn(B, Lq, 2 * Di)
        w, b_ = rn(Di, 4, dtype=torch.float32), rn(Di, dtype=torch.float32)
        g32 = rn(B, Lq, Di, dtype=torch.float32)
        dx = torch.empty((B, Lq, Di), device=dev, dtype=dt)
        dw, db_ = torch.zeros((Di, 4), **f32), torch.zeros(Di, **f32)
        report("conv1d_bwd", timeit(lambda: ops.causal_conv1d_bwd(x[..., :Di], w, b_, g32, dx, dw, db_), flush=flush),
               bytes_=M * Di * (2 * s + 4))
        g2_, g3_ = rn(B, Lq, Di, dtype=torch.float32), rn(B, Lq, Di, dtype=torch.float32)
        report("conv1d_bwd(3 gradient terms, as a Fo-Bi block calls it)",
               timeit(lambda: ops.causal_conv1d_bwd(x[..., :Di], w, b_, g32, dx, dw, db_, dout2=g2_, dout3=g3_), flush=flush),
               bytes_=M * Di * (2 * s + 12))
        if dt != torch.float32:
            h1, h2, h3 = g32.to(dt), g2_.to(dt), g3_.to(dt)
            report("conv1d_bwd(3 gradient terms, 16-bit)",
                   timeit(lambda: ops.causal_conv1d_bwd(x[..., :Di], w, b_, h1, dx, dw, db_, dout2=h2, dout3=h3), flush=flush),
                   bytes_=M * Di * (2 * s + 6))
    if dt != torch.float32 and (not only or "gemm" in only):
        shapes = {"in_proj": (M, 2 * Di, Dm), "out_proj": (M, Dm, Di), "x_proj": (M, R + 2 * N, Di), "dt_proj": (M, Di, R)}
        for name, (m_, n_, k_) in shapes.items():
            kp = (k_ + 7) // 8 * 8
            a, w = rn(m_, kp), rn(n_, kp, sc=k_ ** -0.5)
            odt = torch.float32 if name == "dt_proj" else dt
            out = torch.empty(m_, n_, device=dev, dtype=odt)
            ms = timeit(lambda: ops.gemm_tn(a, w, out=out, k=k_, backend=L.GEMM_TCGEN05), flush=flush)
            by = (m_ * k_ + n_ * k_) * 2 + m_ * n_ * out.element_size()
            report("gemm_" + name, ms, bytes_=by, flops=2.0 * m_ * n_ * k_, shape=[m_, n_, k_])
            if name in ("in_proj", "out_proj"):
                ms = timeit(lambda: torch.matmul(a, w.t(), out=out), flush=flush)
                report("cublas_" + name, ms, flops=2.0 * m_ * n_ * k_)
            if name == "x_proj":      # as the mixer calls it: dt (16-bit, padded pitch) | [B|C] (fp32) split epilogue
                rp = (R + 7) // 8 * 8
                dt_o = torch.empty(m_, rp, device=dev, dtype=dt)
                bc_o = torch.empty(m_, 2 * N, device=dev, dtype=torch.float32)
                ms = timeit(lambda: ops.gemm_tn(a, w, out=dt_o, out2=bc_o, split=R, backend=L.GEMM_TCGEN05), flush=flush)
                report("gemm_x_proj(split dt|BC)", ms, bytes_=(m_ * k_ + n_ * k_) * 2 + m_ * (R * 2 + 2 * N * 4))
            if name == "dt_proj":     # as the mixer calls it: + bias + softplus, fp32 delta
                bias = rn(n_, dtype=torch.float32)
                ms = timeit(lambda: ops.gemm_tn(a, w, out=out, k=k_, bias=bias, act=L.ACT_SOFTPLUS, backend=L.GEMM_TCGEN05), flush=flush)
                report("gemm_dt_proj(+bias+softplus)", ms, bytes_=by)
                out16 = torch.empty(m_, n_, device=dev, dtype=dt)
                ms = timeit(lambda: ops.gemm_tn(a, w, out=out16, k=k_, bias=bias, act=L.ACT_SOFTPLUS, backend=L.GEMM_TCGEN05), flush=flush)
                report("gemm_dt_proj(+bias+softplus, 16-bit delta)", ms, bytes_=(m_ * k_ + n_ * k_) * 2 + m_ * n_ * 2)
            if name == "in_proj":
                ms = timeit(lambda: ops.gemm_tn(a, w, out=out, k=k_, backend=L.GEMM_TCGEN05,
                                                act=L.act_from(L.ACT_SILU, n_ // 2)), flush=flush)
                report("gemm_in_proj+silu(z)", ms, flops=2.0 * m_ * n_ * k_)


if __name__ == "__main__":
    main()
