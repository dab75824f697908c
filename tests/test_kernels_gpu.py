"""Parity of every CUDA kernel against the CPU oracle, through the C ABI (ctypes).  B200 only (-m gpu)."""
import itertools

import pytest
import torch

import aum_oracle as O

pytestmark = pytest.mark.gpu

DEV = "cuda"
DTYPES = [torch.float32, torch.float16, torch.bfloat16]
# comparison of a kernel that computes in fp32 and rounds once to `dt`, against the fp32 oracle fed the SAME
# (already rounded) inputs: a few ulps of the output dtype
TOL = {torch.float32: dict(rtol=2e-5, atol=2e-6),
       torch.float16: dict(rtol=2e-3, atol=2e-3),
       torch.bfloat16: dict(rtol=1.6e-2, atol=1.6e-2)}


def gen(seed):
    return torch.Generator().manual_seed(seed)


def rnd(shape, g, dt=torch.float32, scale=1.0):
    return (scale * torch.randn(shape, generator=g)).to(dt)


# ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("B,Lq,D,W", [(2, 64, 384, 4), (3, 37, 40, 4), (1, 5, 7, 3), (2, 9, 16, 2), (1, 1, 8, 4),
                                      (2, 33, 38, 4),      # D % 4 != 0: two-channel kernel
                                      (1, 513, 1536, 4)])  # one config-2 sequence (4-channel kernel, 33 token tiles)
def test_causal_conv1d(dt, B, Lq, D, W):
    from aum_b200 import ops
    g = gen(1)
    wide = rnd((B, Lq, 2 * D), g, dt)                     # x is the first half of an xz-like buffer
    x = wide[..., :D]
    w = rnd((D, W), g, scale=0.5)
    b = rnd((D,), g, scale=0.5)
    for silu, rev, use_b in itertools.product([True, False], [False, True], [True, False]):
        xc = x.float().permute(0, 2, 1)                   # (B, D, L) for the oracle
        if rev:
            ref = O.causal_conv1d_oracle(xc.flip(-1), w, b if use_b else None, silu).flip(-1)
        else:
            ref = O.causal_conv1d_oracle(xc, w, b if use_b else None, silu)
        out = ops.causal_conv1d(wide.to(DEV)[..., :D], w.to(DEV), b.to(DEV) if use_b else None, silu=silu, reverse=rev)
        torch.testing.assert_close(out.float().cpu().permute(0, 2, 1), ref, **TOL[dt])


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("rows,dim", [(130, 768), (7, 96), (5, 100), (3, 2056)])
def test_add_rmsnorm(dt, rows, dim):
    from aum_b200 import ops
    g = gen(2)
    x = rnd((rows, dim), g, dt)
    r = rnd((rows, dim), g)
    w = 1 + 0.1 * rnd((dim,), g)
    ref_y, ref_r = O.rms_norm_oracle(x.float(), w, None, r, 1e-5, prenorm=True)
    y, res = ops.add_rmsnorm(x.to(DEV), w.to(DEV), None, r.to(DEV), eps=1e-5, prenorm=True)
    assert res.dtype == torch.float32 and y.dtype == dt
    torch.testing.assert_close(res.cpu(), ref_r, rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(y.float().cpu(), ref_y, **TOL[dt])
    # first block: no residual in, fp32 residual out (mamba_models.py:78-87)
    y0, res0 = ops.add_rmsnorm(x.to(DEV), w.to(DEV), None, None, eps=1e-5, prenorm=True)
    ref_y0, ref_r0 = O.rms_norm_oracle(x.float(), w, None, None, 1e-5, prenorm=True)
    torch.testing.assert_close(res0.cpu(), ref_r0, rtol=0, atol=0)
    torch.testing.assert_close(y0.float().cpu(), ref_y0, **TOL[dt])
    # final norm: prenorm=False
    y1 = ops.add_rmsnorm(x.to(DEV), w.to(DEV), None, r.to(DEV), eps=1e-5, prenorm=False)
    torch.testing.assert_close(y1.float().cpu(), ref_y, **TOL[dt])


@pytest.mark.parametrize("sd,dd", [(torch.float32, torch.float32), (torch.float32, torch.bfloat16),
                                   (torch.float16, torch.float16), (torch.bfloat16, torch.float32)])
def test_transpose(sd, dd):
    from aum_b200 import ops
    g = gen(3)
    for B, R, Cc in [(2, 513, 96), (1, 33, 65), (3, 1, 7), (2, 64, 768)]:
        src = rnd((B, R, Cc), g, sd)
        out = ops.transpose(src.to(DEV), dst_dtype=dd)
        assert out.shape == (B, Cc, R)
        torch.testing.assert_close(out.cpu(), src.transpose(1, 2).to(dd), rtol=0, atol=0)


# ----------------------------------------------------------------------------------------------------
def _gemm_ref(a, w, bias=None, row_scale=None, softplus=False):
    acc = a.double() @ w.double().t()
    if row_scale is not None:
        acc = acc * row_scale.double()[:, None]
    if bias is not None:
        acc = acc + bias.double()[None]
    if softplus:
        acc = torch.nn.functional.softplus(acc)
    return acc.float()


@pytest.mark.parametrize("M,N,K", [(70, 50, 33), (300, 256, 128), (64, 80, 192)])
def test_gemm_simt_fp32(M, N, K):
    from aum_b200 import ops, _lib as L
    g = gen(4)
    a, w = rnd((M, K), g), rnd((N, K), g, scale=K ** -0.5)
    bias, rs = rnd((N,), g), 1 + 0.1 * rnd((M,), g)
    out = ops.gemm_tn(a.to(DEV), w.to(DEV), bias=bias.to(DEV), row_scale=rs.to(DEV), act=L.ACT_SOFTPLUS)
    torch.testing.assert_close(out.cpu(), _gemm_ref(a, w, bias, rs, True), rtol=1e-5, atol=1e-5)


# shapes with N >= 512, K >= 256, M >= 1024 and no activation run the CTA-pair (cta_group::2) kernel, the others the
# single-CTA one; AUM_GEMM_PAIR=0 forces the latter everywhere
TC_SHAPES = [
    (300, 256, 128),      # M tail, BN=256
    (128, 64, 64),        # single tile, single k-block
    (1000, 3072, 768),    # in_proj shape (BN=256, 12 k-blocks -> 3 trips round a 4-stage ring)
    (1026, 768, 1536),    # out_proj shape
    (257, 384, 96),       # K tail (96 = 64 + 32), N=384 -> BN=128
    (513, 1536, 48),      # dt_proj: K=48 < one k-block
    (640, 80, 1536),      # x_proj: N=80 -> BN=96
    (100, 24, 200),       # small N -> BN=32, K tail
    (19000, 512, 256),    # more tiles than SMs (persistent loop, both TMEM buffers, phase flips)
    (1100, 600, 320),     # CTA-pair kernel: M, N and K tails (600 = 2 x 256 + 88, 1100 = 4 x 256 + 76)
    (32832, 768, 1536),   # out_proj at config-2 size: 387 pair tiles over 74 pairs
]


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("M,N,K", TC_SHAPES)
def test_gemm_tcgen05(dt, M, N, K):
    from aum_b200 import ops, _lib as L
    g = gen(5)
    a, w = rnd((M, K), g, dt), rnd((N, K), g, dt, scale=K ** -0.5)
    ref = _gemm_ref(a.float(), w.float())
    out = ops.gemm_tn(a.to(DEV), w.to(DEV), out_dtype=torch.float32, backend=L.GEMM_TCGEN05)
    torch.testing.assert_close(out.cpu(), ref, rtol=1e-4, atol=1e-4)
    # same operands through the CUDA-core kernel must agree (independent implementation)
    out_s = ops.gemm_tn(a.to(DEV), w.to(DEV), out_dtype=torch.float32, backend=L.GEMM_SIMT)
    torch.testing.assert_close(out_s.cpu(), ref, rtol=1e-4, atol=1e-4)
    # 16-bit output
    out_h = ops.gemm_tn(a.to(DEV), w.to(DEV), backend=L.GEMM_TCGEN05)
    assert out_h.dtype == dt
    torch.testing.assert_close(out_h.float().cpu(), ref, **TOL[dt])


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
def test_gemm_tcgen05_epilogues(dt):
    """x_proj-style split output, dt_proj-style bias+softplus with a strided/padded A, row scaling."""
    from aum_b200 import ops, _lib as L
    g = gen(6)
    M, Di, R, N = 700, 1536, 48, 16
    u, wx = rnd((M, Di), g, dt), rnd((R + 2 * N, Di), g, dt, scale=Di ** -0.5)
    ref = _gemm_ref(u.float(), wx.float())
    dtb = torch.full((M, 64), 7.0, device=DEV, dtype=dt)            # wider than R: pad columns untouched
    bc = torch.empty((M, 2 * N), device=DEV, dtype=torch.float32)
    ops.gemm_tn(u.to(DEV), wx.to(DEV), out=dtb, out2=bc, split=R, backend=L.GEMM_TCGEN05)
    torch.testing.assert_close(dtb[:, :R].float().cpu(), ref[:, :R], **TOL[dt])
    assert (dtb[:, R:] == 7.0).all()
    torch.testing.assert_close(bc.cpu(), ref[:, R:], rtol=1e-4, atol=1e-4)
    # split not a multiple of 8 (Tiny model: R = 12)
    wx2 = rnd((12 + 2 * N, Di), g, dt, scale=Di ** -0.5)
    dt2 = torch.empty((M, 16), device=DEV, dtype=dt)
    ops.gemm_tn(u.to(DEV), wx2.to(DEV), out=dt2, out2=bc, split=12, backend=L.GEMM_TCGEN05)
    ref2 = _gemm_ref(u.float(), wx2.float())
    torch.testing.assert_close(dt2[:, :12].float().cpu(), ref2[:, :12], **TOL[dt])
    torch.testing.assert_close(bc.cpu(), ref2[:, 12:], rtol=1e-4, atol=1e-4)
    # dt_proj: A = first R columns of a 64-wide buffer, W zero-padded to 64 columns, K = R
    wdt = torch.zeros((Di, 64), dtype=dt)
    wdt[:, :R] = rnd((Di, R), g, dt, scale=R ** -0.5)
    bias = rnd((Di,), g) - 3.0
    dtb.fill_(float("nan"))
    dtb[:, :R] = rnd((M, R), g, dt).to(DEV)
    refd = _gemm_ref(dtb[:, :R].float().cpu(), wdt[:, :R].float(), bias, None, True)
    delta = ops.gemm_tn(dtb, wdt.to(DEV), k=R, bias=bias.to(DEV), act=L.ACT_SOFTPLUS, out_dtype=torch.float32,
                        backend=L.GEMM_TCGEN05)
    torch.testing.assert_close(delta.cpu(), refd, rtol=1e-4, atol=1e-5)
    # the whole softplus range (torch threshold 20, tiny values below -20) and more tiles than SMs through the
    # 16-epilogue-warp kernel: relative error everywhere
    M2 = 20000
    a2 = torch.zeros((M2, 64), dtype=dt)
    a2[:, :R] = rnd((M2, R), g, dt)
    bias2 = torch.linspace(-45.0, 45.0, Di)
    refw = _gemm_ref(a2[:, :R].float(), wdt[:, :R].float(), bias2, None, True)
    dw = ops.gemm_tn(a2.to(DEV), wdt.to(DEV), k=R, bias=bias2.to(DEV), act=L.ACT_SOFTPLUS, out_dtype=torch.float32,
                     backend=L.GEMM_TCGEN05).cpu()
    assert torch.isfinite(dw).all()
    assert ((dw - refw).abs() <= 1e-4 * refw.abs() + 1e-30).all()
    # row scaling
    rs = 1 + 0.2 * rnd((M,), g)
    out = ops.gemm_tn(u.to(DEV), wx.to(DEV), row_scale=rs.to(DEV), out_dtype=torch.float32, backend=L.GEMM_TCGEN05)
    torch.testing.assert_close(out.cpu(), ref * rs[:, None], rtol=1e-4, atol=1e-4)


# ----------------------------------------------------------------------------------------------------
def _scan_inputs(B, Lq, D, N, g, dt):
    u = rnd((B, Lq, D), g, dt)
    delta = rnd((B, Lq, D), g, scale=0.5)
    A = -torch.exp(torch.log(torch.arange(1, N + 1, dtype=torch.float32)).repeat(D, 1) + 0.1 * rnd((D, N), g))
    A_b = -torch.exp(torch.log(torch.arange(1, N + 1, dtype=torch.float32)).repeat(D, 1) + 0.1 * rnd((D, N), g))
    Bm, Cm = rnd((B, Lq, N), g), rnd((B, Lq, N), g)
    Dv = 1 + 0.1 * rnd((D,), g)
    z = rnd((B, Lq, D), g, dt)
    bias = 0.5 * rnd((D,), g) - 2.0
    return u, delta, A, A_b, Bm, Cm, Dv, z, bias


def _oracle_dir(u, delta, A, Bm, Cm, Dv, bias, softplus, reverse):
    """y (+D u) of one direction, token-major in/out, no gate."""
    uc, dc = u.float().permute(0, 2, 1), delta.float().permute(0, 2, 1)
    Bc, Cc = Bm.float().permute(0, 2, 1), Cm.float().permute(0, 2, 1)
    if reverse:
        uc, dc, Bc, Cc = uc.flip(-1), dc.flip(-1), Bc.flip(-1), Cc.flip(-1)
    y = O.selective_scan_oracle(uc, dc, A, Bc, Cc, Dv, None, bias, softplus)
    if reverse:
        y = y.flip(-1)
    return y.permute(0, 2, 1)


SCAN_TOL = {torch.float32: dict(rtol=1e-4, atol=2e-5),
            torch.float16: dict(rtol=4e-3, atol=4e-3),
            torch.bfloat16: dict(rtol=3e-2, atol=3e-2)}


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("B,Lq,D,N", [(2, 64, 384, 16), (2, 37, 40, 16), (1, 1, 8, 16), (1, 19, 16, 8),
                                      (2, 513, 96, 16), (1, 130, 200, 16)])
def test_selective_scan_all_modes(dt, B, Lq, D, N):
    from aum_b200 import ops
    g = gen(7)
    u, delta, A, A_b, Bm, Cm, Dv, z, bias = _scan_inputs(B, Lq, D, N, g, dt)
    cu = lambda t: t.to(DEV)
    yf = _oracle_dir(u, delta, A, Bm, Cm, Dv, bias, True, False)
    yb = _oracle_dir(u, delta, A_b, Bm, Cm, Dv, bias, True, True)
    gate = O.silu_oracle(z.float())

    def mk(Ax, delta_dev=None, bc_dev=None):
        return ops.ScanDirection(cu(u), delta_dev if delta_dev is not None else cu(delta), cu(Ax),
                                 bc_dev[0] if bc_dev else cu(Bm), bc_dev[1] if bc_dev else cu(Cm),
                                 cu(Dv), cu(bias), True)
    # Fo-Bi: both directions in one launch (fp32 delta and B/C, as the fused module passes them)
    out = ops.selective_scan(mk(A), mk(A_b), cu(z))
    torch.testing.assert_close(out.float().cpu(), (yf + yb) * gate, **SCAN_TOL[dt])
    # forward only / backward only / no gate / scale
    out = ops.selective_scan(mk(A), None, cu(z))
    torch.testing.assert_close(out.float().cpu(), yf * gate, **SCAN_TOL[dt])
    out = ops.selective_scan(None, mk(A_b), None, out_scale=0.5)
    torch.testing.assert_close(out.float().cpu(), 0.5 * yb, **SCAN_TOL[dt])
    # delta and B/C in the activation dtype (functional-API mode), interleaved B|C buffer with ld = 2N
    if dt != torch.float32:
        d16 = delta.to(dt)
        bc = torch.cat([Bm, Cm], dim=-1).to(dt).to(DEV)
        yf16 = _oracle_dir(u, d16, A, bc[..., :N].cpu(), bc[..., N:].cpu(), Dv, bias, True, False)
        yb16 = _oracle_dir(u, d16, A_b, bc[..., :N].cpu(), bc[..., N:].cpu(), Dv, bias, True, True)
        out = ops.selective_scan(mk(A, cu(d16), (bc[..., :N], bc[..., N:])), mk(A_b, cu(d16), (bc[..., :N], bc[..., N:])), cu(z))
        torch.testing.assert_close(out.float().cpu(), (yf16 + yb16) * gate, **SCAN_TOL[dt])


def test_selective_scan_plain_options_and_last_state():
    """no D, no z, no bias, no softplus, last_state out (selective_scan_fn's optional outputs)."""
    from aum_b200 import ops
    g = gen(8)
    B, Lq, D, N = 2, 70, 48, 16
    u, delta, A, _, Bm, Cm, _, _, _ = _scan_inputs(B, Lq, D, N, g, torch.float32)
    delta = delta.abs() * 0.2
    ref, last = O.selective_scan_oracle(u.permute(0, 2, 1), delta.permute(0, 2, 1), A, Bm.permute(0, 2, 1),
                                        Cm.permute(0, 2, 1), None, None, None, False, True)
    ls = torch.empty((B, D, N), device=DEV)
    d = ops.ScanDirection(u.to(DEV), delta.to(DEV), A.to(DEV), Bm.to(DEV), Cm.to(DEV), None, None, False, ls)
    out = ops.selective_scan(d, None, None)
    torch.testing.assert_close(out.cpu().permute(0, 2, 1), ref, rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(ls.cpu(), last, rtol=1e-4, atol=2e-5)


def test_selective_scan_properties_at_full_size():
    """BASELINE config-2 scan shape (per sequence: L=513, Di=1536, N=16), batch 8: size-independent properties
    instead of the (too slow) oracle — direction additivity, time-reversal symmetry, linearity in u."""
    from aum_b200 import ops
    g = gen(9)
    B, Lq, D, N = 8, 513, 1536, 16
    u, delta, A, A_b, Bm, Cm, Dv, z, bias = [t.to(DEV) if t is not None else None
                                             for t in _scan_inputs(B, Lq, D, N, g, torch.float32)]
    mk = lambda Ax, uu=u, dd=delta, bb=Bm, cc=Cm: ops.ScanDirection(uu, dd, Ax, bb, cc, Dv, bias, True)
    both = ops.selective_scan(mk(A), mk(A_b), z)
    f = ops.selective_scan(mk(A), None, z)
    b = ops.selective_scan(None, mk(A_b), z)
    torch.testing.assert_close(both, f + b, rtol=1e-5, atol=1e-5)
    # reverse scan of the flipped sequence == flip of the forward scan (same A)
    fl = lambda t: t.flip(1).contiguous()
    b_fl = ops.selective_scan(None, mk(A, fl(u), fl(delta), fl(Bm), fl(Cm)), fl(z))
    torch.testing.assert_close(fl(b_fl), f, rtol=1e-5, atol=1e-5)
    # linear in u for fixed delta/B/C/z
    f2 = ops.selective_scan(mk(A, 2.0 * u), None, z)
    torch.testing.assert_close(f2, 2.0 * f, rtol=1e-5, atol=1e-5)
    assert torch.isfinite(both).all()


@pytest.mark.parametrize("Lq", [2, 7, 8, 9, 15, 16, 17, 1024, 4096])
def test_selective_scan_length_sweep(Lq):
    """BASELINE config 5 lengths (up to 4096) and every tile-boundary case of the TMA ring (L around multiples of 8),
    fused forward+reverse, fp32, against the oracle."""
    from aum_b200 import ops
    g = gen(10 + Lq)
    B, D, N = (1, 64, 16) if Lq > 1000 else (2, 128, 16)
    u, delta, A, A_b, Bm, Cm, Dv, z, bias = _scan_inputs(B, Lq, D, N, g, torch.float32)
    delta = torch.nn.functional.softplus(delta + bias[None, None])      # final delta (module path: no in-kernel bias)
    cu = lambda t: t.to(DEV)
    bc = torch.cat([Bm, Cm], dim=-1).to(DEV).contiguous()               # packed fp32 [B|C] -> TMA-streamed kernel
    mk = lambda Ax: ops.ScanDirection(cu(u), cu(delta), cu(Ax), bc[..., :N], bc[..., N:], cu(Dv))
    out = ops.selective_scan(mk(A), mk(A_b), cu(z))
    yf = _oracle_dir(u, delta, A, Bm, Cm, Dv, None, False, False)
    yb = _oracle_dir(u, delta, A_b, Bm, Cm, Dv, None, False, True)
    torch.testing.assert_close(out.cpu(), (yf + yb) * O.silu_oracle(z), rtol=2e-4, atol=5e-5)
    # same through the generic kernel (interleaved but non-contiguous B / C views force it)
    Bd, Cd = cu(Bm).contiguous(), cu(Cm).contiguous()
    mk2 = lambda Ax: ops.ScanDirection(cu(u), cu(delta), cu(Ax), Bd, Cd, cu(Dv))
    out2 = ops.selective_scan(mk2(A), mk2(A_b), cu(z))
    torch.testing.assert_close(out2.cpu(), out.cpu(), rtol=1e-5, atol=1e-5)


def test_empty_batch_and_zero_length_are_noops():
    from aum_b200 import ops
    w = torch.zeros(8, 4, device=DEV)
    assert ops.causal_conv1d(torch.zeros(0, 5, 8, device=DEV), w, None).shape == (0, 5, 8)
    assert ops.gemm_tn(torch.zeros(0, 16, device=DEV), torch.zeros(4, 16, device=DEV)).shape == (0, 4)
    y = ops.add_rmsnorm(torch.zeros(0, 16, device=DEV), torch.ones(16, device=DEV))
    assert y.shape == (0, 16)


def test_scan_192_channel_build():
    """The one-CTA-per-SM build of the TMA-streamed scan (AUM_SCAN_TMA_CH=192, 6 warps per direction) is selected per
    process, so it runs in a child: every scan parity case again."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, AUM_SCAN_TMA_CH="192")
    r = subprocess.run([sys.executable, "-m", "pytest", __file__, "-q", "-x", "-p", "no:cacheprovider", "-m", "gpu",
                        "-k", "selective_scan_all_modes or length_sweep or properties_at_full_size"],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("dt", DTYPES)
def test_patchify_and_token_assembly(dt):
    """Front end either side of the patch-embed GEMM against the reference formulation: conv2d(img, W, stride=patch)
    .flatten(2).transpose(1, 2) == patchify(x) @ W.flatten(1)^T, and cls insertion + position embedding
    (tokenization.py:278-310,414-451; mamba_models.py:510-541)."""
    from aum_b200 import ops
    g = gen(12)
    B, T_, F_, Dm, p = 3, 64, 32, 24, (16, 16)
    x = rnd((B, T_, F_), g)
    w = rnd((Dm, 1, p[0], p[1]), g, scale=1 / 16)
    cols = ops.patchify(x.to(DEV), p, dt)
    ref_cols = x.view(B, T_ // p[1], p[1], F_ // p[0], p[0]).permute(0, 3, 1, 4, 2).reshape(-1, p[0] * p[1])
    torch.testing.assert_close(cols.float().cpu(), ref_cols.to(dt).float(), rtol=0, atol=0)
    img = x.unsqueeze(1).transpose(2, 3)                                  # (B, 1, F, T)
    ref_tok = torch.nn.functional.conv2d(img, w, None, stride=p).flatten(2).transpose(1, 2)
    tok = ref_cols @ w.reshape(Dm, -1).t()
    torch.testing.assert_close(tok.view(B, -1, Dm), ref_tok, rtol=1e-5, atol=1e-5)
    if dt == torch.float32:
        N = ref_tok.shape[1]
        pos, cls = rnd((1, N + 1, Dm), g), rnd((1, 1, Dm), g)
        tp = N // 2
        ref = torch.cat((ref_tok[:, :tp], cls.expand(B, -1, -1), ref_tok[:, tp:]), dim=1) \
            + torch.cat((pos[:, 1:tp + 1], pos[:, :1], pos[:, tp + 1:]), dim=1)
        out = ops.assemble_tokens(ref_tok.contiguous().to(DEV), pos[0].contiguous().to(DEV), cls.reshape(-1).to(DEV))
        torch.testing.assert_close(out.cpu(), ref, rtol=0, atol=0)


# ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("reverse", [False, True])
@pytest.mark.parametrize("B,Lq,Di,R,N", [(2, 513, 1536, 48, 16),      # AuM-Base: 80 outputs -> 96-wide MMA, 24 channel blocks
                                         (3, 65, 768, 24, 16),        # AuM-Small: 56 outputs -> 64-wide MMA; tiles span sequences
                                         (1, 7, 64, 8, 16),           # one short sequence, one channel block
                                         (5, 130, 200, 16, 16)])      # Di not a multiple of 64 (channel tail), 650 tokens
def test_fused_conv_xproj(dt, reverse, B, Lq, Di, R, N):
    """aum_conv_xproj_fwd (conv + SiLU as the producer of x_proj's tensor-core operand) against the oracle's conv and a
    float64 x_proj on the SAME rounded u, and against the two separate kernels it replaces."""
    from aum_b200 import ops
    g = gen(40 + Di)
    xz = rnd((B, Lq, 2 * Di), g, dt)                      # x = first half of an xz buffer (z must never be read)
    w = rnd((Di, 4), g, scale=0.5)
    b = rnd((Di,), g, scale=0.5)
    wx = rnd((R + 2 * N, Di), g, dt, scale=Di ** -0.5)
    x = xz.to(DEV)[..., :Di]
    M = B * Lq
    dtb = torch.full((M, R + 8), 7.0, device=DEV, dtype=dt)
    bc = torch.empty((M, 2 * N), device=DEV, dtype=torch.float32)
    assert ops.conv_xproj_eligible(x, w.to(DEV), wx.to(DEV), R, 2 * N)
    u = ops.conv_xproj(x, w.to(DEV), b.to(DEV), wx.to(DEV), R, dtb, bc, reverse=reverse)
    # conv vs oracle
    xc = xz[..., :Di].float().permute(0, 2, 1)
    ref_u = (O.causal_conv1d_oracle(xc.flip(-1), w, b, True).flip(-1) if reverse else O.causal_conv1d_oracle(xc, w, b, True))
    torch.testing.assert_close(u.float().cpu().permute(0, 2, 1), ref_u, **TOL[dt])
    # x_proj of the u that was actually produced (exactly representable operands -> only fp32 accumulation differs)
    ref_x = _gemm_ref(u.float().cpu().view(M, Di), wx.float())
    torch.testing.assert_close(dtb[:, :R].float().cpu(), ref_x[:, :R], **TOL[dt])
    assert (dtb[:, R:] == 7.0).all()
    torch.testing.assert_close(bc.cpu(), ref_x[:, R:], rtol=1e-4, atol=1e-4)
    # the two kernels it replaces
    u2 = ops.causal_conv1d(x, w.to(DEV), b.to(DEV), silu=True, reverse=reverse)
    torch.testing.assert_close(u.float(), u2.float(), **TOL[dt])
