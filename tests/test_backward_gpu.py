"""Gradients of the training path (row a12 of SURVEY.md section 8a: BiMambaInnerFn.backward and friends) against
torch autograd through the CPU oracle.  B200 only (-m gpu)."""
import pytest
import torch

import aum_oracle as O
from conftest import load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda"


def gen(seed):
    return torch.Generator().manual_seed(seed)


def rnd(shape, g, scale=1.0):
    return scale * torch.randn(shape, generator=g)


def _close(a, b, rtol, atol, msg=""):
    scale = max(b.abs().max().item(), 1e-12)
    torch.testing.assert_close(a.float().cpu(), b.float(), rtol=rtol, atol=atol * max(scale, 1.0), msg=lambda m: f"{msg}: {m}")


@pytest.mark.parametrize("reverse", [False, True])
@pytest.mark.parametrize("B,Lq,D,W", [(2, 37, 40, 4), (1, 130, 6, 3), (2, 16, 64, 4)])
def test_causal_conv1d_bwd(reverse, B, Lq, D, W):
    from aum_b200 import ops
    g = gen(1)
    x = rnd((B, Lq, D), g).requires_grad_()
    w = rnd((D, W), g, 0.5).requires_grad_()
    b = rnd((D,), g, 0.5).requires_grad_()
    G = rnd((B, Lq, D), g)
    xc = x.permute(0, 2, 1)
    y = O.causal_conv1d_oracle(xc.flip(-1) if reverse else xc, w, b, True)
    y = (y.flip(-1) if reverse else y).permute(0, 2, 1)
    (y * G).sum().backward()
    dx = torch.empty((B, Lq, D), device=DEV)
    dw = torch.zeros((D, W), device=DEV)
    db = torch.zeros((D,), device=DEV)
    ops.causal_conv1d_bwd(x.detach().to(DEV), w.detach().to(DEV), b.detach().to(DEV), G.to(DEV), dx, dw, db,
                          silu=True, reverse=reverse)
    _close(dx, x.grad, 1e-4, 1e-5, "dx")
    _close(dw, w.grad, 1e-4, 1e-5, "dw")
    _close(db, b.grad, 1e-4, 1e-5, "dbias")


@pytest.mark.parametrize("reverse", [False, True])
@pytest.mark.parametrize("B,Lq,D,W,terms", [(2, 513, 128, 4, 3), (1, 1, 64, 4, 1), (3, 2, 66, 4, 2), (2, 70, 34, 2, 3),
                                            (1, 1030, 64, 4, 2), (2, 9, 7, 4, 3)])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_causal_conv1d_bwd_streaming_kernel_multi_term(reverse, B, Lq, D, W, terms, dt):
    """The production conv backward (segment-walking kernel): x as a strided view of a wider buffer (the x half of xz),
    1-3 gradient terms summed on the fly (scan du of both directions + the x_proj term), segment boundaries inside the
    sequence (L = 513 / 1030), sequences shorter than the 3-position halo, an odd channel count (tile-kernel fallback);
    against autograd through the oracle on the same (rounded) inputs."""
    from aum_b200 import ops
    g = gen(3)
    xw = rnd((B, Lq, 2 * D + 2), g).to(dt)
    x = xw[..., :D].float().clone().requires_grad_()
    w = rnd((D, W), g, 0.5).requires_grad_()
    b = rnd((D,), g, 0.5).requires_grad_()
    Gs = [rnd((B, Lq, D), g) for _ in range(terms)]
    xc = x.permute(0, 2, 1)
    y = O.causal_conv1d_oracle(xc.flip(-1) if reverse else xc, w, b, True)
    y = (y.flip(-1) if reverse else y).permute(0, 2, 1)
    (y * sum(Gs)).sum().backward()
    xd = xw.to(DEV)[..., :D]
    dx = torch.empty((B, Lq, D), device=DEV, dtype=dt)
    dw = torch.zeros((D, W), device=DEV)
    db = torch.zeros((D,), device=DEV)
    Gd = [G_.to(DEV) for G_ in Gs] + [None, None]
    ops.causal_conv1d_bwd(xd, w.detach().to(DEV), b.detach().to(DEV), Gd[0], dx, dw, db, silu=True, reverse=reverse,
                          dout2=Gd[1], dout3=Gd[2])
    tol = (1e-4, 1e-5) if dt == torch.float32 else (2e-2, 1e-2)
    _close(dx, x.grad, *tol, "dx")
    _close(dw, w.grad, 2e-4, 2e-5, "dw")
    _close(db, b.grad, 2e-4, 2e-5, "dbias")


def test_causal_conv1d_bwd_tile_kernel_still_agrees(monkeypatch):
    """AUM_CONV_BWD_TILE=1 (read once per process) selects the tile kernel: run it in a child process against the
    streaming kernel's result."""
    import os, subprocess, sys, textwrap
    code = textwrap.dedent("""
        import sys, torch
        sys.path.insert(0, %r)
        from aum_b200 import ops
        g = torch.Generator().manual_seed(4)
        x = torch.randn(2, 100, 64, generator=g).cuda(); w = torch.randn(64, 4, generator=g).cuda(); b = torch.randn(64, generator=g).cuda()
        G = torch.randn(2, 100, 64, generator=g).cuda()
        dx = torch.empty_like(x); dw = torch.zeros(64, 4, device="cuda"); db = torch.zeros(64, device="cuda")
        ops.causal_conv1d_bwd(x, w, b, G, dx, dw, db, silu=True)
        torch.save({"dx": dx.cpu(), "dw": dw.cpu(), "db": db.cpu()}, sys.argv[1])
    """) % os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "audio-mamba-aum_b200")
    import tempfile
    outs = []
    for env_extra in ({}, {"AUM_CONV_BWD_TILE": "1"}):
        with tempfile.NamedTemporaryFile(suffix=".pt") as f:
            subprocess.run([sys.executable, "-c", code, f.name], check=True, env=dict(os.environ, **env_extra))
            outs.append(torch.load(f.name))
    for k in ("dx", "dw", "db"):
        torch.testing.assert_close(outs[0][k], outs[1][k], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("dt", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("rows,dim,prenorm,has_res", [(130, 768, True, True), (37, 96, True, False), (20, 384, False, True),
                                                      (9, 100, True, True)])
def test_add_rmsnorm_backward(dt, rows, dim, prenorm, has_res):
    """rms_norm_fn under autograd (native aum_add_rmsnorm_bwd; dim=100 takes the generic formula path)."""
    from mamba_ssm.ops.triton.layernorm import rms_norm_fn
    g = gen(3)
    x = rnd((rows, dim), g).to(dt)
    res = rnd((rows, dim), g) if has_res else None
    w = 1 + 0.1 * rnd((dim,), g)
    Gy, Gr = rnd((rows, dim), g), rnd((rows, dim), g)
    xr = x.float().clone().requires_grad_()
    rr = res.clone().requires_grad_() if has_res else None
    wr = w.clone().requires_grad_()
    yo, ro = O.rms_norm_oracle(xr, wr, None, rr, 1e-5, prenorm=True)
    loss = (yo * Gy).sum() + ((ro * Gr).sum() if prenorm else 0)
    loss.backward()
    xd = x.to(DEV).requires_grad_()
    rd = res.to(DEV).requires_grad_() if has_res else None
    wd = w.to(DEV).requires_grad_()
    out = rms_norm_fn(xd, wd, None, residual=rd, prenorm=prenorm, residual_in_fp32=True, eps=1e-5)
    if prenorm:
        ((out[0].float() * Gy.to(DEV)).sum() + (out[1] * Gr.to(DEV)).sum()).backward()
    else:
        (out.float() * Gy.to(DEV)).sum().backward()
    tol = {torch.float32: 1e-4, torch.float16: 4e-3, torch.bfloat16: 3e-2}[dt]
    _close(xd.grad, xr.grad, tol, tol, "dx")
    _close(wd.grad, wr.grad, max(tol, 1e-3), tol, "dweight")
    if has_res:
        _close(rd.grad, rr.grad, tol, tol, "dresidual")


def _scan_ref(u, delta, A, A_b, Bm, Cm, Dv, z, scale, dirs):
    """token-major oracle: out = scale * (y_f + y_b) * silu(z) with differentiable torch ops."""
    uc, dc = u.permute(0, 2, 1), delta.permute(0, 2, 1)
    Bc, Cc = Bm.permute(0, 2, 1), Cm.permute(0, 2, 1)
    y = 0
    if "f" in dirs:
        y = y + O.selective_scan_oracle(uc, dc, A, Bc, Cc, Dv, None, None, False)
    if "b" in dirs:
        y = y + O.selective_scan_oracle(uc.flip(-1), dc.flip(-1), A_b, Bc.flip(-1), Cc.flip(-1), Dv, None, None, False).flip(-1)
    y = y.permute(0, 2, 1)
    return scale * y * O.silu_oracle(z), y


@pytest.mark.parametrize("dirs", ["fb", "f", "b"])
@pytest.mark.parametrize("B,Lq,D", [(2, 37, 40), (1, 8, 64), (2, 70, 96), (1, 1, 16)])
def test_selective_scan_bwd_vs_oracle_autograd(dirs, B, Lq, D):
    from aum_b200 import ops
    N = 16
    g = gen(2)
    u = rnd((B, Lq, D), g).requires_grad_()
    delta = (0.05 + 0.3 * torch.rand((B, Lq, D), generator=g)).requires_grad_()
    A = (-torch.exp(torch.log(torch.arange(1, N + 1.0)).repeat(D, 1) + 0.1 * rnd((D, N), g))).requires_grad_()
    A_b = (-torch.exp(torch.log(torch.arange(1, N + 1.0)).repeat(D, 1) + 0.1 * rnd((D, N), g))).requires_grad_()
    Bm, Cm = rnd((B, Lq, N), g).requires_grad_(), rnd((B, Lq, N), g).requires_grad_()
    Dv = (1 + 0.1 * rnd((D,), g)).requires_grad_()
    z = rnd((B, Lq, D), g).requires_grad_()
    G = rnd((B, Lq, D), g)
    scale = 0.5
    out, ypre = _scan_ref(u, delta, A, A_b, Bm, Cm, Dv, z, scale, dirs)
    (out * G).sum().backward()

    cu = lambda t: t.detach().to(DEV).contiguous()
    bc = torch.cat([cu(Bm), cu(Cm)], dim=-1).contiguous()
    du = torch.empty((B, Lq, D), device=DEV); dd = torch.empty_like(du)
    dbc = torch.zeros((B, Lq, 2 * N), device=DEV)
    dA = torch.zeros((D, N), device=DEV); dAb = torch.zeros((D, N), device=DEV); dD = torch.zeros((D,), device=DEV)
    dz = torch.empty((B, Lq, D), device=DEV); oz = torch.empty_like(dz)
    mk = lambda Ax, dAx: ops.ScanBwdDirection(cu(u), cu(delta), cu(Ax), bc, cu(Dv), du, dd, dAx, dD, dbc,
                                              ops.scan_bwd_workspace(B, Lq, D, DEV))
    ops.selective_scan_bwd(mk(A, dA) if "f" in dirs else None, mk(A_b, dAb) if "b" in dirs else None,
                           cu(z), cu(ypre), G.to(DEV), dz, oz, out_scale=scale)
    tol = dict(rtol=2e-4, atol=2e-5)
    _close(oz, out.detach(), msg="out_z", **tol)
    _close(dz, z.grad, msg="dz", **tol)
    _close(du, u.grad, msg="du", **tol)
    _close(dd, delta.grad, msg="ddelta", **tol)
    _close(dbc[..., :N], Bm.grad, msg="dB", **tol)
    _close(dbc[..., N:], Cm.grad, msg="dC", **tol)
    _close(dD, Dv.grad, msg="dD", **tol)
    if "f" in dirs:
        _close(dA, A.grad, msg="dA", **tol)
    if "b" in dirs:
        _close(dAb, A_b.grad, msg="dA_b", **tol)


@pytest.mark.parametrize("generic", [False, True])
@pytest.mark.parametrize("dirs", ["fb", "f", "b"])
@pytest.mark.parametrize("B,Lq,D,dt", [(2, 37, 40, torch.float32), (1, 8, 64, torch.float32), (2, 70, 96, torch.float32),
                                       (1, 1, 16, torch.float32), (2, 513, 192, torch.float32),
                                       (2, 130, 128, torch.float16), (2, 130, 128, torch.bfloat16)])
def test_selective_scan_bwd_from_forward_checkpoints(dirs, B, Lq, D, dt, generic, monkeypatch):
    """The training pair as the autograd function uses it: the forward kernel leaves state checkpoints and the
    pre-gate output, the backward kernel (TMA-streamed, or generic with AUM_SCAN_BWD_GENERIC) consumes them.
    Fo-Bi sharing: both directions accumulate into the same du / ddelta."""
    from aum_b200 import ops
    if generic:
        monkeypatch.setenv("AUM_SCAN_BWD_GENERIC", "1")
    N = 16
    g = gen(3)
    q = lambda t: t.to(dt).float()                      # values exactly representable in the activation dtype
    u = q(rnd((B, Lq, D), g)).requires_grad_()
    delta = (0.05 + 0.3 * torch.rand((B, Lq, D), generator=g)).requires_grad_()
    A = (-torch.exp(torch.log(torch.arange(1, N + 1.0)).repeat(D, 1) + 0.1 * rnd((D, N), g))).requires_grad_()
    A_b = (-torch.exp(torch.log(torch.arange(1, N + 1.0)).repeat(D, 1) + 0.1 * rnd((D, N), g))).requires_grad_()
    Bm, Cm = rnd((B, Lq, N), g).requires_grad_(), rnd((B, Lq, N), g).requires_grad_()
    Dv = (1 + 0.1 * rnd((D,), g)).requires_grad_()
    z = q(rnd((B, Lq, D), g)).requires_grad_()
    G = q(rnd((B, Lq, D), g))
    scale = 0.5 if dirs == "fb" else 1.0
    out, ypre_ref = _scan_ref(u, delta, A, A_b, Bm, Cm, Dv, z, scale, dirs)
    (out * G).sum().backward()

    cu = lambda t: t.detach().to(DEV).contiguous()
    ud, zd, dl = cu(u).to(dt), cu(z).to(dt), cu(delta)
    bc = torch.cat([cu(Bm), cu(Cm)], dim=-1).contiguous()
    ck = {k: ops.scan_bwd_workspace(B, Lq, D, DEV) for k in "fb"}
    mkf = lambda Ax, k: ops.ScanDirection(ud, dl, cu(Ax), bc[..., :N], bc[..., N:], cu(Dv), ckpt=ck[k])
    y_pre = torch.empty((B, Lq, D), device=DEV, dtype=dt)
    out_d = ops.selective_scan(mkf(A, "f") if "f" in dirs else None, mkf(A_b, "b") if "b" in dirs else None, zd,
                               out_scale=scale, y_pre=y_pre)
    lo = dt != torch.float32
    ftol = dict(rtol=2e-2, atol=2e-2) if lo else dict(rtol=2e-4, atol=2e-5)
    _close(out_d.float(), out.detach(), msg="out", **ftol)

    du = torch.full((B, Lq, D), float("nan"), device=DEV); dd = torch.full_like(du, float("nan"))
    dbc = torch.zeros((B, Lq, 2 * N), device=DEV)
    dA = torch.zeros((D, N), device=DEV); dAb = torch.zeros((D, N), device=DEV); dD = torch.zeros((D,), device=DEV)
    dz = torch.empty((B, Lq, D), device=DEV, dtype=dt); oz = torch.empty_like(dz)
    mk = lambda Ax, dAx, k: ops.ScanBwdDirection(ud, dl, cu(Ax), bc, cu(Dv), du, dd, dAx, dD, dbc, ck[k], ckpt_valid=True)
    ops.selective_scan_bwd(mk(A, dA, "f") if "f" in dirs else None, mk(A_b, dAb, "b") if "b" in dirs else None,
                           zd, y_pre, G.to(DEV).to(dt), dz, oz, out_scale=scale)
    # 16-bit tiers: y_pre is stored rounded, so dz / out_z carry that rounding; the fp32 outputs do not
    tol = dict(rtol=2e-2, atol=2e-2) if lo else dict(rtol=2e-4, atol=2e-5)
    gt = dict(rtol=2e-4, atol=5e-5)
    _close(oz.float(), out.detach(), msg="out_z", **tol)
    _close(dz.float(), z.grad, msg="dz", **tol)
    _close(du, u.grad, msg="du", **gt)
    _close(dd, delta.grad, msg="ddelta", **gt)
    _close(dbc[..., :N], Bm.grad, msg="dB", **gt)
    _close(dbc[..., N:], Cm.grad, msg="dC", **gt)
    _close(dD, Dv.grad, msg="dD", **gt)
    if "f" in dirs:
        _close(dA, A.grad, msg="dA", **gt)
    if "b" in dirs:
        _close(dAb, A_b.grad, msg="dA_b", **gt)


@pytest.mark.parametrize("bt,kw", [("v1", {}), ("none", {}), ("v2", {"if_devide_out": True}), ("v1", {"bias": True, "init_layer_scale": 0.5})])
def test_mamba_module_backward_vs_oracle_autograd(bt, kw):
    """All parameter gradients and d(hidden) of one mixer, fp32 tier, against autograd through the oracle."""
    from mamba_ssm.modules.mamba_simple import Mamba
    Dm, Lq, B = 64, 33, 2
    torch.manual_seed(5)
    m = Mamba(Dm, bimamba_type=bt, **kw)
    g = gen(6)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith(("A_log", "A_b_log", "D", "D_b")):
                p.add_(0.1 * torch.randn(p.shape, generator=g))
            if n.endswith("in_proj.bias") or n.endswith("out_proj.bias"):
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    hidden = rnd((B, Lq, Dm), g)
    G = rnd((B, Lq, Dm), g)
    # oracle autograd on CPU
    ref_p = {k: v.detach().clone().requires_grad_() for k, v in m.state_dict().items()}
    h_ref = hidden.clone().requires_grad_()
    out_ref = O.mamba_forward_oracle(ref_p, h_ref, bt, kw.get("if_devide_out", False))
    (out_ref * G).sum().backward()
    # engine
    m = m.to(DEV)
    h = hidden.to(DEV).requires_grad_()
    out = m(h)
    torch.testing.assert_close(out.detach().cpu(), out_ref.detach(), rtol=1e-4, atol=1e-5)
    (out * G.to(DEV)).sum().backward()
    _close(h.grad, h_ref.grad, 5e-4, 5e-5, "d hidden")
    for n, p in m.named_parameters():
        assert p.grad is not None, n
        _close(p.grad, ref_p[n].grad, 1e-3, 1e-4, n)


def test_mamba_module_backward_with_generic_scan_bwd_kernel(monkeypatch):
    """Same gradients through the generic (non-TMA) backward kernel."""
    monkeypatch.setenv("AUM_SCAN_BWD_GENERIC", "1")
    test_mamba_module_backward_vs_oracle_autograd("v1", {})
    test_mamba_module_backward_vs_oracle_autograd("v2", {"if_devide_out": True})


def test_mamba_module_backward_with_generic_scan_kernel(monkeypatch):
    """Same gradients when the forward runs the generic (non-TMA) scan kernel: its checkpoints must line up too."""
    monkeypatch.setenv("AUM_SCAN_GENERIC", "1")
    test_mamba_module_backward_vs_oracle_autograd("v1", {})
    test_mamba_module_backward_vs_oracle_autograd("v2", {"if_devide_out": True})


@pytest.mark.parametrize("bt", ["v1", "v2", "none"])
@pytest.mark.parametrize("dt,budget", [(torch.float16, 2e-2), (torch.bfloat16, 8e-2)])
def test_mamba_module_backward_16bit_budget(dt, budget, bt):
    """16-bit training tiers (d_inner = 256: the specialised backward scan with 16-bit delta / du / ddelta, the 16-bit
    gradient terms of the conv backward - three for Fo-Bi, two for Bi-Bi and Fo-Fo) against fp32 autograd through the oracle."""
    from mamba_ssm.modules.mamba_simple import Mamba
    Dm, Lq, B = 128, 65, 2
    p = O.make_mamba_params(Dm, bimamba_type=bt, seed=31, perturb_A=0.1)
    g = gen(7)
    hidden = rnd((B, Lq, Dm), g)
    G = rnd((B, Lq, Dm), g)
    ref_p = {k: v.clone().requires_grad_() for k, v in p.items()}
    h_ref = hidden.clone().requires_grad_()
    (O.mamba_forward_oracle(ref_p, h_ref, bt) * G).sum().backward()
    m = Mamba(Dm, bimamba_type=bt).to(DEV)
    m.load_state_dict(p)
    h = hidden.to(DEV).to(dt).requires_grad_()
    (m(h).float() * G.to(DEV)).sum().backward()
    rel = lambda a, b: (a.float().cpu() - b).abs().max().item() / max(b.abs().max().item(), 1e-9)
    assert rel(h.grad, h_ref.grad) < budget
    for n, q in m.named_parameters():
        assert rel(q.grad, ref_p[n].grad) < budget, n


def test_audio_mamba_training_step_matches_oracle_autograd():
    """A tiny Fo-Bi AudioMamba: loss.backward() through the engine vs autograd through the oracle (fp32 tier)."""
    from aum_b200.audio_mamba import AudioMamba
    c = load_golden("audio_mamba_tiny.pt")["fobi_tiny"]
    kw = c["kwargs"]
    sd = c["state"]
    x = c["x"]
    tgt = (torch.rand(x.shape[0], kw["num_classes"], generator=gen(8)) > 0.7).float()
    ref_p = {k: v.clone().requires_grad_() for k, v in sd.items()}
    logits_ref = O.audio_mamba_forward_oracle(ref_p, x, depth=kw["depth"], bimamba_type="v1")
    torch.nn.functional.binary_cross_entropy_with_logits(logits_ref, tgt).backward()
    m = AudioMamba(**kw).to(DEV)
    m.load_state_dict(sd, strict=True)
    logits = m(x.to(DEV))
    torch.testing.assert_close(logits.detach().cpu(), logits_ref.detach(), rtol=1e-3, atol=1e-5)
    torch.nn.functional.binary_cross_entropy_with_logits(logits, tgt.to(DEV)).backward()
    for n, p in m.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), n
        ref = ref_p[n].grad
        err = (p.grad.cpu() - ref).abs().max().item()
        assert err <= 2e-3 * max(ref.abs().max().item(), 1e-6) + 1e-7, (n, err, ref.abs().max().item())


# ----------------------------------------------------------------------------------------------------
# The reference's functional ops are autograd.Functions (selective_scan_interface.py:14-74, 155-289, 292-434, 437-603):
# the shim's are differentiable too, with the reference's (batch, channel, length) layouts.
def _fn_params(Dm, g, bt="v1"):
    p = O.make_mamba_params(Dm, bimamba_type=bt, seed=77, perturb_A=0.1)
    return {k: v.clone() for k, v in p.items()}


@pytest.mark.parametrize("which", ["bimamba_inner_fn", "mamba_inner_fn", "mamba_inner_fn_no_out_proj"])
@pytest.mark.parametrize("channel_major", [True, False])
def test_functional_inner_ops_backward_vs_oracle_autograd(which, channel_major):
    import mamba_ssm.ops.selective_scan_interface as F_
    Dm, Lq, B = 64, 37, 2
    Di = 2 * Dm
    g = gen(21)
    p = _fn_params(Dm, g)
    xz0 = rnd((B, 2 * Di, Lq), g)
    G = rnd((B, Lq, Dm), g) if which != "mamba_inner_fn_no_out_proj" else rnd((B, Di, Lq), g)
    names = ["conv1d.weight", "conv1d.bias", "x_proj.weight", "dt_proj.weight", "out_proj.weight", "A_log", "A_b_log", "D",
             "dt_proj.bias"]

    def run(dev, fns):
        q = {k: p[k].clone().to(dev).requires_grad_() for k in names}
        if channel_major or dev == "cpu":
            xz = xz0.clone().to(dev).requires_grad_()
            xz_arg = xz
        else:       # a transposed view of a token-major buffer (what this package's own module passes)
            xz = xz0.transpose(1, 2).contiguous().to(dev).requires_grad_()
            xz_arg = xz.transpose(1, 2)
        A, A_b = -torch.exp(q["A_log"]), -torch.exp(q["A_b_log"])
        if which == "bimamba_inner_fn":
            out = fns[0](xz_arg, q["conv1d.weight"], q["conv1d.bias"], q["x_proj.weight"], q["dt_proj.weight"],
                         q["out_proj.weight"], None, A, A_b, None, None, q["D"], q["dt_proj.bias"])
        elif which == "mamba_inner_fn":
            out = fns[1](xz_arg, q["conv1d.weight"], q["conv1d.bias"], q["x_proj.weight"], q["dt_proj.weight"],
                         q["out_proj.weight"], None, A, None, None, q["D"], q["dt_proj.bias"])
        else:
            out = fns[2](xz_arg, q["conv1d.weight"], q["conv1d.bias"], q["x_proj.weight"], q["dt_proj.weight"],
                         A, None, None, q["D"], q["dt_proj.bias"])
        (out * G.to(dev)).sum().backward()
        gx = xz.grad if (channel_major or dev == "cpu") else xz.grad.transpose(1, 2)
        return out.detach().cpu(), gx.cpu(), {k: (v.grad.cpu() if v.grad is not None else None) for k, v in q.items()}

    oracle = (lambda *a: O.bimamba_inner_oracle(*a[:7], a[7], a[8], a[11], a[12]),
              lambda *a: O.mamba_inner_oracle(*a[:7], a[7], a[10], a[11]),
              lambda *a: O.mamba_inner_no_out_proj_oracle(*a[:5], a[5], a[8], a[9]))
    out_r, gx_r, gr = run("cpu", oracle)
    out_d, gx_d, gd = run(DEV, (F_.bimamba_inner_fn, F_.mamba_inner_fn, F_.mamba_inner_fn_no_out_proj))
    torch.testing.assert_close(out_d, out_r, rtol=1e-4, atol=1e-5)
    _close(gx_d, gx_r, 5e-4, 5e-5, "dxz")
    for k in names:
        if gr[k] is None:
            assert gd[k] is None or gd[k].abs().max() == 0, k
            continue
        _close(gd[k], gr[k], 1e-3, 1e-4, k)


def test_functional_selective_scan_fn_and_conv_backward_vs_oracle_autograd():
    import mamba_ssm.ops.selective_scan_interface as F_
    from causal_conv1d import causal_conv1d_fn
    g = gen(22)
    B, D, Lq, N = 2, 48, 29, 16
    base = dict(u=rnd((B, D, Lq), g), delta=0.5 * rnd((B, D, Lq), g),
                A=-torch.exp(torch.log(torch.arange(1, N + 1.0)).repeat(D, 1) + 0.1 * rnd((D, N), g)),
                Bm=rnd((B, N, Lq), g), Cm=rnd((B, N, Lq), g), Dv=1 + 0.1 * rnd((D,), g), z=rnd((B, D, Lq), g),
                bias=0.5 * rnd((D,), g) - 1.0)
    G = rnd((B, D, Lq), g)

    def run(dev, fn, use_z):
        q = {k: v.clone().to(dev).requires_grad_() for k, v in base.items()}
        out = fn(q["u"], q["delta"], q["A"], q["Bm"], q["Cm"], q["Dv"], q["z"] if use_z else None, q["bias"], True)
        (out * G.to(dev)).sum().backward()
        return out.detach().cpu(), {k: (v.grad.cpu() if v.grad is not None else None) for k, v in q.items()}

    for use_z in (True, False):
        out_r, gr = run("cpu", lambda u, d, A, Bm, Cm, Dv, z, b, sp: O.selective_scan_oracle(u, d, A, Bm, Cm, Dv, z, b, sp), use_z)
        out_d, gd = run(DEV, F_.selective_scan_fn, use_z)
        torch.testing.assert_close(out_d, out_r, rtol=1e-4, atol=2e-5)
        for k in base:
            if k == "z" and not use_z:
                continue
            _close(gd[k], gr[k], 5e-4, 5e-5, f"scan d{k} (z={use_z})")
    # causal_conv1d_fn
    x, w, b = rnd((B, D, Lq), g), rnd((D, 4), g, 0.5), rnd((D,), g, 0.5)
    xr, wr, br = (t.clone().requires_grad_() for t in (x, w, b))
    (O.causal_conv1d_oracle(xr, wr, br, True) * G).sum().backward()
    xd, wd, bd = (t.clone().to(DEV).requires_grad_() for t in (x, w, b))
    (causal_conv1d_fn(xd, wd, bd, "silu") * G.to(DEV)).sum().backward()
    _close(xd.grad, xr.grad, 1e-4, 1e-5, "conv dx")
    _close(wd.grad, wr.grad, 1e-4, 1e-5, "conv dw")
    _close(bd.grad, br.grad, 1e-4, 1e-5, "conv db")


@pytest.mark.parametrize("embed,act,bt,generic", [(192, torch.bfloat16, "v1", False), (96, torch.float32, "v2", False),
                                                  (64, torch.float32, "v1", True)])
def test_direct_gradient_accumulation_into_a_flat_buffer(embed, act, bt, generic, monkeypatch):
    """FlatGradReducer marks parameters for in-place gradient accumulation: the kernels write straight into the flat
    buffer (no temporaries, no AccumulateGrad; dA * A lands in the A_log gradient from inside the backward scan -
    specialised, general and generic kernel) and the result equals the autograd-delivered gradients."""
    from aum_b200.audio_mamba import AudioMamba
    from aum_b200.dist import FlatGradReducer
    if generic:
        monkeypatch.setenv("AUM_SCAN_BWD_GENERIC", "1")
    torch.manual_seed(11)
    kw = dict(embed_dim=embed, depth=2, num_classes=35, spectrogram_size=(128, 128), bimamba_type=bt, act_dtype=act)
    a = AudioMamba(**kw).to(DEV)
    b = AudioMamba(**kw).to(DEV)
    b.load_state_dict(a.state_dict())
    x = 0.5 * torch.randn(3, 128, 128, device=DEV)
    tgt = (torch.rand(3, 35, device=DEV) > 0.7).float()
    torch.nn.functional.binary_cross_entropy_with_logits(a(x), tgt).backward()
    red = FlatGradReducer(b.parameters())
    red.zero()
    torch.nn.functional.binary_cross_entropy_with_logits(b(x), tgt).backward()
    for (n, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
        assert pb.grad.data_ptr() >= red.flat.data_ptr() and pb.grad.data_ptr() < red.flat.data_ptr() + red.flat.numel() * 4, n
        scale = max(pa.grad.abs().max().item(), 1e-9)
        assert (pa.grad - pb.grad).abs().max().item() <= (2e-2 if act != torch.float32 else 1e-4) * scale, (n, scale)


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16, torch.float16])
def test_sum_cast_colsum(dt):
    """aum_sum_cast_colsum: (a + b) cast to the activation dtype and its column sums in one pass (dt_proj chain of the
    backward, selective_scan_interface.py:556,583-586); a third gradient term of the conv backward."""
    from aum_b200 import ops
    g = gen(31)
    for rows, cols in [(513, 1536), (70, 128), (1, 4), (130, 768)]:
        a, b = rnd((rows, cols), g), rnd((rows, cols), g)
        cs = torch.full((cols,), 2.0, device=DEV)
        out = ops.sum_cast_colsum(a.to(DEV), b.to(DEV), dt, cs)
        torch.testing.assert_close(out.cpu(), (a + b).to(dt), rtol=0, atol=0)
        torch.testing.assert_close(cs.cpu(), 2.0 + (a + b).sum(0), rtol=1e-5, atol=1e-4)
        out1 = ops.sum_cast_colsum(a.to(DEV), None, dt)
        torch.testing.assert_close(out1.cpu(), a.to(dt), rtol=0, atol=0)
    # conv backward with three gradient terms == with their sum
    B, Lq, D = 2, 37, 64
    x, w, bias = rnd((B, Lq, D), g), rnd((D, 4), g, 0.5), rnd((D,), g, 0.5)
    g1, g2, g3 = rnd((B, Lq, D), g), rnd((B, Lq, D), g), rnd((B, Lq, D), g)
    res = []
    for terms in ((g1 + g2 + g3, None, None), (g1, g2, g3)):
        dx = torch.empty((B, Lq, D), device=DEV); dw = torch.zeros((D, 4), device=DEV); db = torch.zeros((D,), device=DEV)
        ops.causal_conv1d_bwd(x.to(DEV), w.to(DEV), bias.to(DEV), terms[0].to(DEV), dx, dw, db,
                              dout2=terms[1].to(DEV) if terms[1] is not None else None,
                              dout3=terms[2].to(DEV) if terms[2] is not None else None)
        res.append((dx.cpu(), dw.cpu(), db.cpu()))
    for p_, q_ in zip(res[0], res[1]):
        torch.testing.assert_close(p_, q_, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
def test_16bit_gradient_terms_equal_the_rounded_fp32_ones(dt):
    """The training step keeps du / ddelta of the backward scan and the gradient terms of the conv backward in the
    activation dtype (aum_scan_bwd_dir.dgrad_dtype, aum_causal_conv1d_bwd(dout_dtype), aum_sum_cast_colsum(in_dtype)):
    the 16-bit outputs are the fp32 ones rounded once, and consumers fed the rounded terms agree with fp32 arithmetic on
    those same rounded values."""
    from aum_b200 import ops
    g = gen(41)
    B, Lq, D, N = 2, 70, 256, 16
    u, z, G = (rnd((B, Lq, D), g).to(DEV).to(dt) for _ in range(3))
    delta = torch.nn.functional.softplus(rnd((B, Lq, D), g) - 2.0).to(DEV)
    bc = rnd((B, Lq, 2 * N), g).to(DEV)
    A = -torch.exp(rnd((D, N), g, 0.3)).to(DEV); A_b = -torch.exp(rnd((D, N), g, 0.3)).to(DEV)
    Dv = torch.ones(D, device=DEV)
    ck = {k: ops.scan_bwd_workspace(B, Lq, D, DEV) for k in "fb"}
    y_pre = torch.empty((B, Lq, D), device=DEV, dtype=dt)
    mkf = lambda Ax, k: ops.ScanDirection(u, delta, Ax, bc[..., :N], bc[..., N:], Dv, ckpt=ck[k])
    ops.selective_scan(mkf(A, "f"), mkf(A_b, "b"), z, y_pre=y_pre)
    res = {}
    for gdt in (torch.float32, dt):
        du = [torch.full((B, Lq, D), float("nan"), device=DEV, dtype=gdt) for _ in range(2)]
        dd = [torch.full((B, Lq, D), float("nan"), device=DEV, dtype=gdt) for _ in range(2)]
        dbc = torch.zeros((B, Lq, 2 * N), device=DEV)
        dA = [torch.zeros((D, N), device=DEV) for _ in range(2)]
        dD = torch.zeros((D,), device=DEV)
        dz = torch.empty((B, Lq, D), device=DEV, dtype=dt); oz = torch.empty_like(dz)
        mk = lambda Ax, i, k: ops.ScanBwdDirection(u, delta, Ax, bc, Dv, du[i], dd[i], dA[i], dD, dbc, ck[k], ckpt_valid=True)
        ops.selective_scan_bwd(mk(A, 0, "f"), mk(A_b, 1, "b"), z, y_pre, G, dz, oz, softplus_grad=True)
        res[gdt] = (du, dd, dbc, dA, dz)
    f32r, lo = res[torch.float32], res[dt]
    for i in range(2):
        assert torch.equal(lo[0][i], f32r[0][i].to(dt)) and torch.equal(lo[1][i], f32r[1][i].to(dt))
    torch.testing.assert_close(lo[2], f32r[2], rtol=1e-5, atol=1e-5)          # dB|dC, dA, dz do not depend on the store dtype
    torch.testing.assert_close(lo[3][0], f32r[3][0], rtol=1e-5, atol=1e-5)
    assert torch.equal(lo[4], f32r[4])
    # consumers: the 16-bit terms summed in fp32
    du_f, du_b = lo[0]
    g2 = rnd((B, Lq, D), g).to(DEV).to(dt)
    x = rnd((B, Lq, D), g).to(DEV).to(dt); w = rnd((D, 4), g, 0.5).to(DEV); bias = rnd((D,), g, 0.5).to(DEV)
    out = []
    for terms in ((du_f, g2, du_b), (du_f.float(), g2.float(), du_b.float())):
        dx = torch.empty((B, Lq, D), device=DEV, dtype=dt); dw = torch.zeros((D, 4), device=DEV); db = torch.zeros((D,), device=DEV)
        ops.causal_conv1d_bwd(x, w, bias, terms[0], dx, dw, db, dout2=terms[1], dout3=terms[2])
        out.append((dx, dw, db))
    assert torch.equal(out[0][0], out[1][0])
    torch.testing.assert_close(out[0][1], out[1][1], rtol=1e-5, atol=1e-4)
    torch.testing.assert_close(out[0][2], out[1][2], rtol=1e-5, atol=1e-4)
    dd_f, dd_b = lo[1]
    cs0, cs1 = torch.zeros(D, device=DEV), torch.zeros(D, device=DEV)
    s0 = ops.sum_cast_colsum(dd_f.view(-1, D), dd_b.view(-1, D), dt, cs0)
    s1 = ops.sum_cast_colsum(dd_f.float().view(-1, D), dd_b.float().view(-1, D), dt, cs1)
    assert torch.equal(s0, s1)
    torch.testing.assert_close(cs0, cs1, rtol=1e-5, atol=1e-4)
    # delta in the activation dtype as well (the training default): same results as fp32 arithmetic on the rounded delta
    d16 = delta.to(dt)
    outs = []
    for dl_ in (d16, d16.float()):
        ckd = {k: ops.scan_bwd_workspace(B, Lq, D, DEV) for k in "fb"}
        yp = torch.empty((B, Lq, D), device=DEV, dtype=dt)
        mkf2 = lambda Ax, k: ops.ScanDirection(u, dl_, Ax, bc[..., :N], bc[..., N:], Dv, ckpt=ckd[k])
        o = ops.selective_scan(mkf2(A, "f"), mkf2(A_b, "b"), z, y_pre=yp)
        gd = dt if dl_.dtype == dt else torch.float32
        du = [torch.empty((B, Lq, D), device=DEV, dtype=gd) for _ in range(2)]
        dd = [torch.empty((B, Lq, D), device=DEV, dtype=gd) for _ in range(2)]
        dbc = torch.zeros((B, Lq, 2 * N), device=DEV)
        dA = [torch.zeros((D, N), device=DEV) for _ in range(2)]
        dD = torch.zeros((D,), device=DEV)
        dz = torch.empty((B, Lq, D), device=DEV, dtype=dt); oz = torch.empty_like(dz)
        mk = lambda Ax, i, k: ops.ScanBwdDirection(u, dl_, Ax, bc, Dv, du[i], dd[i], dA[i], dD, dbc, ckd[k], ckpt_valid=True)
        ops.selective_scan_bwd(mk(A, 0, "f"), mk(A_b, 1, "b"), z, yp, G, dz, oz, softplus_grad=True)
        outs.append((o, [t.to(dt) for t in du], [t.to(dt) for t in dd], dbc, dA[0], dz))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][5], outs[1][5])
    for i in range(2):
        assert torch.equal(outs[0][1][i], outs[1][1][i]) and torch.equal(outs[0][2][i], outs[1][2][i])
    torch.testing.assert_close(outs[0][3], outs[1][3], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(outs[0][4], outs[1][4], rtol=1e-5, atol=1e-5)
    # outside the specialised configuration the 16-bit form is refused, loudly
    du16 = torch.empty((B, Lq, D), device=DEV, dtype=dt)
    with pytest.raises(Exception):
        ops.selective_scan_bwd(ops.ScanBwdDirection(u, delta, A, bc, Dv, du16, du16.clone(), dA[0], dD, dbc, ck["f"], ckpt_valid=True),
                               None, None, None, G, None, None)
