"""Parity of the reference-facing API (Mamba module, functional ops, RMSNorm) against the golden vectors produced
by the real reference and against the CPU oracle.  B200 only (-m gpu)."""
import pytest
import torch

import aum_oracle as O
from conftest import load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda"

# fp32 tier: the SURVEY's op-level tolerance vs the reference's *_ref functions
F32 = dict(rtol=1e-4, atol=1e-5)


def _rel_err(out, ref):
    return (out.float().cpu() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-6)


def test_selective_scan_fn_matches_reference_golden():
    """selective_scan_fn with the reference's (B, D, L) layout and option set (selective_scan_interface.py:77-83)."""
    from mamba_ssm.ops.selective_scan_interface import selective_scan_fn
    cu = lambda t: t.to(DEV) if isinstance(t, torch.Tensor) else t
    for name, c in load_golden("selective_scan_ref.pt").items():
        res = selective_scan_fn(cu(c["u"]), cu(c["delta"]), cu(c["A"]), cu(c["B"]), cu(c["C"]), cu(c["D"]),
                                cu(c["z"]), cu(c["delta_bias"]), c["delta_softplus"], c["return_last_state"])
        if c["return_last_state"]:
            out, last = res
            torch.testing.assert_close(last.cpu(), c["last_state"], msg=name, **F32)
        else:
            out = res
        assert out.shape == c["out"].shape, name
        torch.testing.assert_close(out.cpu(), c["out"], msg=name, **F32)


def test_inner_fns_match_reference_golden_cfg1():
    """BASELINE config 1 (d_model=192, d_state=16, L=64): bimamba_inner_fn / mamba_inner_fn /
    mamba_inner_fn_no_out_proj on the reference's channel-major xz, fp32 tier."""
    from mamba_ssm.ops.selective_scan_interface import bimamba_inner_fn, mamba_inner_fn, mamba_inner_fn_no_out_proj
    c = {k: (v.to(DEV) if isinstance(v, torch.Tensor) else v) for k, v in load_golden("inner_ref_cfg1.pt").items()}
    common = (c["xz"], c["conv1d_weight"], c["conv1d_bias"], c["x_proj_weight"], c["delta_proj_weight"],
              c["out_proj_weight"])
    with torch.no_grad():
        out = bimamba_inner_fn(*common, None, c["A"], c["A_b"], None, None, c["D"], delta_bias=c["delta_bias"],
                               delta_softplus=True)
        torch.testing.assert_close(out, c["out_bimamba"], **F32)
        out = bimamba_inner_fn(*common, c["out_proj_bias"], c["A"], c["A_b"], None, None, c["D"],
                               delta_bias=c["delta_bias"], delta_softplus=True)
        torch.testing.assert_close(out, c["out_bimamba_bias"], **F32)
        out = mamba_inner_fn(*common, None, c["A"], None, None, c["D"], delta_bias=c["delta_bias"], delta_softplus=True)
        torch.testing.assert_close(out, c["out_mamba"], **F32)
        # no_out_proj returns (B, Di, L); projecting it must reproduce mamba_inner_ref
        y = mamba_inner_fn_no_out_proj(*common[:5], c["A"], None, None, c["D"], delta_bias=c["delta_bias"],
                                       delta_softplus=True)
        assert y.shape == (c["xz"].shape[0], c["xz"].shape[1] // 2, c["xz"].shape[2])
        proj = torch.nn.functional.linear(y.transpose(1, 2), c["out_proj_weight"])
        torch.testing.assert_close(proj, c["out_mamba"], rtol=2e-4, atol=2e-5)


@pytest.mark.parametrize("case", ["v1", "v2_divide", "v2_nodivide", "none", "v1_gamma_bias"])
def test_mamba_module_matches_reference_golden(case):
    """Mamba(...).forward for Fo-Bi / Bi-Bi / Fo-Fo loaded with the reference module's own state dict."""
    from mamba_ssm.modules.mamba_simple import Mamba
    c = load_golden("mamba_module.pt")[case]
    m = Mamba(c["d_model"], **c["kwargs"]).to(DEV)
    m.load_state_dict(c["state"], strict=True)
    with torch.no_grad():
        out = m(c["hidden"].to(DEV))
    assert out.dtype == torch.float32 and out.shape == c["out"].shape
    torch.testing.assert_close(out.cpu(), c["out"], **F32)
    # 16-bit tiers (tcgen05 GEMMs): error budget relative to the output scale
    with torch.no_grad():
        e16 = _rel_err(m(c["hidden"].to(DEV).half()), c["out"])
        eb16 = _rel_err(m(c["hidden"].to(DEV).bfloat16()), c["out"])
        with torch.autocast("cuda", dtype=torch.float16):
            oa = m(c["hidden"].to(DEV))
    assert oa.dtype == torch.float16
    assert e16 < 1e-2, e16
    assert eb16 < 6e-2, eb16
    assert _rel_err(oa, c["out"]) < 1e-2


@pytest.mark.parametrize("dt,budget", [(torch.float32, 2e-5), (torch.float16, 4e-3), (torch.bfloat16, 3e-2)])
def test_mamba_block_base_size_vs_oracle(dt, budget):
    """One AuM-Base Fo-Bi block (Dm=768, Di=1536, R=48, N=16, L=513) on 2 sequences against the CPU oracle."""
    from mamba_ssm.modules.mamba_simple import Mamba
    p = O.make_mamba_params(768, bimamba_type="v1", seed=11, perturb_A=0.1)
    g = torch.Generator().manual_seed(12)
    hidden = torch.randn(2, 513, 768, generator=g)
    ref = O.mamba_forward_oracle(p, hidden, "v1")
    m = Mamba(768, bimamba_type="v1").to(DEV)
    m.load_state_dict(p, strict=True)
    with torch.no_grad():
        out = m(hidden.to(DEV).to(dt))
    assert _rel_err(out, ref) < budget


def test_rmsnorm_module_matches_reference_golden():
    from mamba_ssm.ops.triton.layernorm import RMSNorm, rms_norm_fn
    c = load_golden("rms_norm_ref.pt")
    with torch.no_grad():
        out, res = rms_norm_fn(c["x"].to(DEV), c["weight"].to(DEV), None, residual=c["residual"].to(DEV),
                               prenorm=True, residual_in_fp32=True, eps=c["eps"])
        torch.testing.assert_close(out.cpu(), c["out"], rtol=2e-5, atol=2e-6)
        torch.testing.assert_close(res.cpu(), c["residual_out"], rtol=1e-6, atol=1e-6)
        n = RMSNorm(c["x"].shape[-1], eps=c["eps"]).to(DEV)
        n.weight.copy_(c["weight"])
        torch.testing.assert_close(n(c["x"].to(DEV)).cpu(), c["out_nores"], rtol=2e-5, atol=2e-6)


def test_functional_ops_refuse_modes_outside_the_aum_path_loudly():
    """The functional ops are differentiable (tests/test_backward_gpu.py); what they do not implement they refuse loudly
    instead of falling back: complex A, grouped B/C, and a differentiable scan with d_state != 16."""
    from mamba_ssm.ops.selective_scan_interface import selective_scan_fn
    u = torch.randn(1, 8, 16, device=DEV)
    Bm = torch.randn(1, 16, 16, device=DEV)
    with pytest.raises(NotImplementedError):
        selective_scan_fn(u, u, -torch.ones(8, 16, device=DEV, dtype=torch.complex64), Bm, Bm)
    with pytest.raises(NotImplementedError):
        selective_scan_fn(u, u, -torch.ones(8, 16, device=DEV), torch.randn(1, 2, 16, 16, device=DEV), Bm)
    with pytest.raises(NotImplementedError):
        selective_scan_fn(u.clone().requires_grad_(), u, -torch.ones(8, 8, device=DEV), Bm[:, :8], Bm[:, :8])
