"""Parity of the BENCHMARKED precision tier (fp16 / bf16 activations on the tcgen05 + TMA kernels) by the criterion
SURVEY.md section 8(c) states, plus vs-oracle cases at the full config-2 / config-5 scan shapes.  B200 only (-m gpu).

Criterion (16-bit tiers).  Three results on the same inputs:
    ref  = the fp32 oracle (restatement of selective_scan_ref / bimamba_inner_ref / AudioMamba.forward),
    emu  = the *emulated reference*: the same oracle with every tensor the reference materialises in the autocast
           dtype rounded to it (oracle ``io_dtype=``; rounding points selective_scan_interface.py:452-468,499-507,517),
    out  = this repo's CUDA path in that dtype (through the C ABI),
and the kernel path must be no further from the fp32 oracle than twice the reference's own half-precision
arithmetic is:   err(out, ref) <= 2 * err(emu, ref) + atol,   for the max-abs and the rms error, atol = 1e-4 of the
output scale.  The fp32 tier keeps the north star's rtol 1e-3.  Every measured error is written to
``gpurun_out/r2_parity.json`` (committed copy: profiles/r2_parity.json).
"""
import json
import os

import pytest
import torch

import aum_oracle as O
from conftest import ROOT

pytestmark = pytest.mark.gpu
DEV = "cuda"
N = 16


def _record(key, val):
    path = os.path.join(ROOT, "gpurun_out", "r2_parity.json")
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        d = json.load(open(path)) if os.path.isfile(path) else {}
        d[key] = val
        json.dump(d, open(path, "w"), indent=1, sort_keys=True)
    except OSError:
        pass


def _errs(a, ref):
    scale = ref.abs().max().item()
    d = (a.double() - ref.double())
    return {"max_rel_to_scale": d.abs().max().item() / scale, "rms_rel_to_scale": d.pow(2).mean().sqrt().item() / scale}


def _check_tier(name, dt, out, emu, ref):
    ek, ee = _errs(out, ref), _errs(emu, ref)
    rec = {"dtype": str(dt).replace("torch.", ""), "scale_max_abs_ref": ref.abs().max().item(),
           "kernel_vs_fp32_oracle": ek, "emulated_reference_vs_fp32_oracle": ee,
           "criterion": "err_kernel <= 2*err_emulated + 1e-4 (relative to max|ref|)"}
    _record(name, rec)
    print(name, json.dumps(rec))
    for k in ("max_rel_to_scale", "rms_rel_to_scale"):
        assert ek[k] <= 2.0 * ee[k] + 1e-4, (name, k, ek, ee)


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("bt", ["v1", "v2"])
def test_block_16bit_tier_vs_emulated_reference(dt, bt):
    """One AuM-Base mixer block at the config-2 sequence shape (Dm=768, Di=1536, L=513), batch 2."""
    from aum_b200.modules import Mamba
    p = O.make_mamba_params(768, bimamba_type=bt, seed=41, perturb_A=0.1)
    g = torch.Generator().manual_seed(42)
    hidden = torch.randn(2, 513, 768, generator=g)
    ref = O.mamba_forward_oracle(p, hidden, bt, if_devide_out=True)
    emu = O.mamba_forward_oracle(p, hidden, bt, if_devide_out=True, io_dtype=dt)
    m = Mamba(768, bimamba_type=bt, if_devide_out=True).to(DEV)
    m.load_state_dict(p, strict=True)
    with torch.no_grad():
        out = m(hidden.to(DEV).to(dt)).float().cpu()
        out32 = m(hidden.to(DEV)).float().cpu()
    torch.testing.assert_close(out32, ref, rtol=1e-3, atol=1e-4 * ref.abs().max().item())      # fp32 tier
    _record(f"block_{bt}_fp32", {"kernel_vs_fp32_oracle": _errs(out32, ref)})
    _check_tier(f"block_{bt}_{str(dt).replace('torch.', '')}", dt, out, emu, ref)


def test_aum_base_model_16bit_tiers_vs_emulated_reference():
    """BASELINE config 2 (AuM-Base Fo-Bi, depth 24, 527 classes, 128x1024 mel), 2 clips: logits of the fp32, fp16
    (the benchmarked dtype) and bf16 tiers."""
    from aum_b200.audio_mamba import AudioMamba
    sd = O.make_audio_mamba_state(768, 24, num_classes=527, seed=21, perturb_A=0.1)
    x = O.make_spectrogram(2, (128, 1024), seed=22)
    ref = O.audio_mamba_forward_oracle(sd, x, depth=24, bimamba_type="v1")
    m = AudioMamba(embed_dim=768, depth=24, num_classes=527, bimamba_type="v1").to(DEV).eval()
    m.load_state_dict(sd, strict=True)
    with torch.no_grad():
        out32 = m(x.to(DEV)).cpu()
    torch.testing.assert_close(out32, ref, rtol=1e-3, atol=1e-3 * ref.abs().max().item())       # north star, fp32 tier
    _record("model_config2_fp32", {"kernel_vs_fp32_oracle": _errs(out32, ref), "rtol": 1e-3})
    for dt in (torch.float16, torch.bfloat16):
        emu = O.audio_mamba_forward_oracle(sd, x, depth=24, bimamba_type="v1", io_dtype=dt)
        m.act_dtype = dt
        with torch.no_grad():
            out = m(x.to(DEV)).cpu()
        _check_tier(f"model_config2_{str(dt).replace('torch.', '')}", dt, out, emu, ref)
        assert (out.argmax(-1) == ref.argmax(-1)).all()
    # the bench path itself: CUDA graph + two sequence groups, fp16, same clips
    m2 = AudioMamba(embed_dim=768, depth=24, num_classes=527, bimamba_type="v1", act_dtype=torch.float16,
                    use_cuda_graph=True, micro_batches=2).to(DEV).eval()
    m2.load_state_dict(sd, strict=True)
    x4 = torch.cat([x, x], dim=0)
    with torch.no_grad():
        o4 = m2(x4.pin_memory()).cpu()
    emu = O.audio_mamba_forward_oracle(sd, x, depth=24, bimamba_type="v1", io_dtype=torch.float16)
    _check_tier("model_config2_fp16_graph_2groups", torch.float16, o4[:2], emu, ref)
    assert torch.equal(o4[:2], o4[2:])


def test_aum_small_bibi_model_16bit_tier_vs_emulated_reference():
    """BASELINE config 4 model (AuM-Small Bi-Bi, depth 24), 1 clip."""
    from aum_b200.audio_mamba import AudioMamba
    sd = O.make_audio_mamba_state(384, 24, num_classes=527, bimamba_type="v2", seed=31, perturb_A=0.1)
    x = O.make_spectrogram(1, (128, 1024), seed=32)
    ref = O.audio_mamba_forward_oracle(sd, x, depth=24, bimamba_type="v2")
    m = AudioMamba(embed_dim=384, depth=24, num_classes=527, bimamba_type="v2").to(DEV).eval()
    m.load_state_dict(sd, strict=True)
    for dt in (torch.float16, torch.bfloat16):
        emu = O.audio_mamba_forward_oracle(sd, x, depth=24, bimamba_type="v2", io_dtype=dt)
        m.act_dtype = dt
        with torch.no_grad():
            out = m(x.to(DEV)).cpu()
        _check_tier(f"model_config4_{str(dt).replace('torch.', '')}", dt, out, emu, ref)


# ----------------------------------------------------------------------------------------------------
def _scan_case(B, Lq, D, g, dt):
    u = torch.randn((B, Lq, D), generator=g).to(dt)
    z = torch.randn((B, Lq, D), generator=g).to(dt)
    bias = 0.5 * torch.randn((D,), generator=g) - 2.0
    delta = torch.nn.functional.softplus(0.5 * torch.randn((B, Lq, D), generator=g) + bias)
    mkA = lambda: -torch.exp(torch.log(torch.arange(1, N + 1, dtype=torch.float32)).repeat(D, 1)
                             + 0.1 * torch.randn((D, N), generator=g))
    A, A_b = mkA(), mkA()
    Bm, Cm = torch.randn((B, Lq, N), generator=g), torch.randn((B, Lq, N), generator=g)
    Dv = 1 + 0.1 * torch.randn((D,), generator=g)
    return u, z, delta, A, A_b, Bm, Cm, Dv


def _oracle_bidir(u, z, delta, A, A_b, Bm, Cm, Dv, gate=None):
    uc, dc = u.float().permute(0, 2, 1), delta.permute(0, 2, 1)
    Bc, Cc = Bm.permute(0, 2, 1), Cm.permute(0, 2, 1)
    yf = O.selective_scan_oracle(uc, dc, A, Bc, Cc, Dv)
    yb = O.selective_scan_oracle(uc.flip(-1), dc.flip(-1), A_b, Bc.flip(-1), Cc.flip(-1), Dv).flip(-1)
    return ((yf + yb).permute(0, 2, 1)) * (O.silu_oracle(z.float()) if gate is None else gate)


@pytest.mark.parametrize("dt", [torch.float32, torch.float16])
@pytest.mark.parametrize("Lq", [513, 4096])
def test_selective_scan_full_width_vs_oracle(Lq, dt):
    """The fused forward+reverse TMA-streamed scan at the FULL channel width of AuM-Base (Di = 1536: 12 CTA columns)
    for one config-2 sequence (L = 513) and one config-5 sequence (L = 4096), against the oracle."""
    from aum_b200 import ops
    g = torch.Generator().manual_seed(900 + Lq)
    u, z, delta, A, A_b, Bm, Cm, Dv = _scan_case(1, Lq, 1536, g, dt)
    ref = _oracle_bidir(u, z, delta, A, A_b, Bm, Cm, Dv)
    cu = lambda t: t.to(DEV)
    bc = torch.cat([Bm, Cm], dim=-1).to(DEV).contiguous()
    mk = lambda Ax: ops.ScanDirection(cu(u), cu(delta), cu(Ax), bc[..., :N], bc[..., N:], cu(Dv))
    out = ops.selective_scan(mk(A), mk(A_b), cu(z)).float().cpu()
    tol = dict(rtol=2e-4, atol=5e-5) if dt == torch.float32 else dict(rtol=4e-3, atol=4e-3)
    torch.testing.assert_close(out, ref, **tol)
    _record(f"scan_fwd_L{Lq}_D1536_{str(dt).replace('torch.', '')}", {"kernel_vs_fp32_oracle": _errs(out, ref)})
    # pre-gated z (what the mixer passes: the in_proj epilogue has applied SiLU)
    zg = O.silu_oracle(z.float()).to(dt)
    out2 = ops.selective_scan(mk(A), mk(A_b), cu(zg), z_pregated=True).float().cpu()
    ref2 = _oracle_bidir(u, z, delta, A, A_b, Bm, Cm, Dv, gate=zg.float())
    torch.testing.assert_close(out2, ref2, **tol)


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_selective_scan_bwd_full_width_vs_oracle_autograd(dt):
    """Training pair at the config-3 sequence shape (L = 513, Di = 1536): forward with checkpoints, then the
    TMA/TMEM backward kernel over 12 CTA columns x 2 directions, against torch autograd through the oracle."""
    from aum_b200 import ops
    B, Lq, D = 1, 513, 1536
    g = torch.Generator().manual_seed(77)
    u, z, delta, A, A_b, Bm, Cm, Dv = _scan_case(B, Lq, D, g, dt)
    G = torch.randn((B, Lq, D), generator=g).to(dt).float()
    leaves = [t.float().clone().requires_grad_() for t in (u, z, delta, A, A_b, Bm, Cm, Dv)]
    ur, zr, dr, Ar, Abr, Br, Cr, Dr = leaves
    out = _oracle_bidir(ur, zr, dr, Ar, Abr, Br, Cr, Dr)
    (out * G).sum().backward()

    cu = lambda t: t.detach().to(DEV).contiguous()
    ud, zd, dl = cu(u), cu(z), cu(delta)
    bc = torch.cat([cu(Bm), cu(Cm)], dim=-1).contiguous()
    ck = {k: ops.scan_bwd_workspace(B, Lq, D, DEV) for k in "fb"}
    y_pre = torch.empty((B, Lq, D), device=DEV, dtype=dt)
    mkf = lambda Ax, k: ops.ScanDirection(ud, dl, cu(Ax), bc[..., :N], bc[..., N:], cu(Dv), ckpt=ck[k])
    out_d = ops.selective_scan(mkf(A, "f"), mkf(A_b, "b"), zd, y_pre=y_pre)
    du = torch.full((B, Lq, D), float("nan"), device=DEV); dd = torch.full_like(du, float("nan"))
    dbc = torch.zeros((B, Lq, 2 * N), device=DEV)
    dA = torch.zeros((D, N), device=DEV); dAb = torch.zeros((D, N), device=DEV); dD = torch.zeros((D,), device=DEV)
    dz = torch.empty((B, Lq, D), device=DEV, dtype=dt); oz = torch.empty_like(dz)
    mk = lambda Ax, dAx, k: ops.ScanBwdDirection(ud, dl, cu(Ax), bc, cu(Dv), du, dd, dAx, dD, dbc, ck[k], ckpt_valid=True)
    ops.selective_scan_bwd(mk(A, dA, "f"), mk(A_b, dAb, "b"), zd, y_pre, G.to(DEV).to(dt), dz, oz)
    lo = dt != torch.float32

    def close(a, b, name, loose=False):
        scale = max(b.abs().max().item(), 1.0)
        rt, at = (3e-2, 3e-2) if (lo and loose) else (3e-4, 1e-4)
        torch.testing.assert_close(a.float().cpu(), b.float(), rtol=rt, atol=at * scale, msg=lambda m: f"{name}: {m}")
        _record(f"scan_bwd_L513_D1536_{str(dt).replace('torch.', '')}_{name}", _errs(a.float().cpu(), b.float()))
    close(out_d, out.detach(), "out", loose=True)
    close(dz, zr.grad, "dz", loose=True)        # y_pre is stored rounded in the 16-bit tier
    close(du, ur.grad, "du")
    close(dd, dr.grad, "ddelta")
    close(dbc[..., :N], Br.grad, "dB")
    close(dbc[..., N:], Cr.grad, "dC")
    close(dD, Dr.grad, "dD")
    close(dA, Ar.grad, "dA")
    close(dAb, Abr.grad, "dA_b")
