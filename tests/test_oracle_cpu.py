"""Pin the CPU oracle (oracle/aum_oracle.py) to golden vectors produced by the real reference
(oracle/make_golden.py).  No GPU."""
import torch

import aum_oracle as O
from conftest import load_golden

TOL = dict(rtol=2e-5, atol=2e-6)


def test_selective_scan_matches_reference_golden():
    cases = load_golden("selective_scan_ref.pt")
    assert len(cases) >= 5
    for name, c in cases.items():
        res = O.selective_scan_oracle(c["u"], c["delta"], c["A"], c["B"], c["C"], c["D"], c["z"],
                                      c["delta_bias"], c["delta_softplus"], c["return_last_state"])
        if c["return_last_state"]:
            out, last = res
            torch.testing.assert_close(last, c["last_state"], **TOL, msg=name)
        else:
            out = res
        torch.testing.assert_close(out, c["out"], **TOL, msg=name)


def test_selective_scan_fp64_bounds_fp32_rounding():
    c = load_golden("selective_scan_ref.pt")["cfg1_scan"]
    o64 = O.selective_scan_oracle(c["u"], c["delta"], c["A"], c["B"], c["C"], c["D"], c["z"],
                                  c["delta_bias"], True, compute_dtype=torch.float64)
    assert (o64 - c["out"]).abs().max() < 5e-5


def test_inner_ops_match_reference_golden():
    c = load_golden("inner_ref_cfg1.pt")
    common = (c["xz"], c["conv1d_weight"], c["conv1d_bias"], c["x_proj_weight"], c["delta_proj_weight"],
              c["out_proj_weight"])
    out = O.bimamba_inner_oracle(*common, None, c["A"], c["A_b"], c["D"], c["delta_bias"])
    torch.testing.assert_close(out, c["out_bimamba"], rtol=1e-4, atol=1e-5)
    out = O.bimamba_inner_oracle(*common, c["out_proj_bias"], c["A"], c["A_b"], c["D"], c["delta_bias"])
    torch.testing.assert_close(out, c["out_bimamba_bias"], rtol=1e-4, atol=1e-5)
    out = O.mamba_inner_oracle(*common, None, c["A"], c["D"], c["delta_bias"])
    torch.testing.assert_close(out, c["out_mamba"], rtol=1e-4, atol=1e-5)


def test_mamba_module_forward_matches_reference_golden():
    cases = load_golden("mamba_module.pt")
    assert set(cases) >= {"v1", "v2_divide", "v2_nodivide", "none", "v1_gamma_bias"}
    for name, c in cases.items():
        kw = c["kwargs"]
        out = O.mamba_forward_oracle(c["state"], c["hidden"], kw.get("bimamba_type", "none"),
                                     kw.get("if_devide_out", False))
        torch.testing.assert_close(out, c["out"], rtol=1e-4, atol=1e-5, msg=name)


def test_rms_norm_matches_reference_golden():
    c = load_golden("rms_norm_ref.pt")
    out, res = O.rms_norm_oracle(c["x"], c["weight"], None, c["residual"], c["eps"], prenorm=True)
    torch.testing.assert_close(out, c["out"], **TOL)
    torch.testing.assert_close(res, c["residual_out"], **TOL)
    out0 = O.rms_norm_oracle(c["x"], c["weight"], None, None, c["eps"], prenorm=False)
    torch.testing.assert_close(out0, c["out_nores"], **TOL)


def test_audio_mamba_forward_matches_reference_golden():
    cases = load_golden("audio_mamba_tiny.pt")
    for name, c in cases.items():
        kw = c["kwargs"]
        logits = O.audio_mamba_forward_oracle(c["state"], c["x"], depth=kw["depth"],
                                              bimamba_type=kw["bimamba_type"], if_devide_out=True)
        torch.testing.assert_close(logits, c["logits"], rtol=1e-4, atol=1e-5, msg=name)
        feats = O.audio_mamba_forward_oracle(c["state"], c["x"], depth=kw["depth"],
                                             bimamba_type=kw["bimamba_type"], if_devide_out=True,
                                             return_features=True)
        torch.testing.assert_close(feats, c["features"], rtol=1e-4, atol=1e-5, msg=name)


def test_param_generators_have_reference_shapes():
    # AuM-Base Fo-Bi with 527 classes must reproduce the README's 92.1 M
    sd = O.make_audio_mamba_state(768, 24, num_classes=527)
    n = sum(v.numel() for v in sd.values())
    assert n == 92_107_535
    sd = O.make_audio_mamba_state(384, 24, num_classes=527, bimamba_type="v2")
    assert sum(v.numel() for v in sd.values()) == 25_539_215


def test_golden_vectors_regenerate_from_the_real_reference(tmp_path):
    """The pin itself: where the reference tree is present (the build container), oracle/make_golden.py re-runs the
    REAL reference functions and must reproduce every committed tensor of tests/golden/*.pt.  Skipped on the GPU box
    (no /root/reference there)."""
    import os
    import subprocess
    import sys
    import pytest
    if not os.path.isdir("/root/reference/vim-mamba_ssm"):
        pytest.skip("reference tree not present")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, AUM_GOLDEN_OUT=str(tmp_path))
    r = subprocess.run([sys.executable, os.path.join(root, "oracle", "make_golden.py")], env=env, capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]

    def same(a, b, path):
        if isinstance(a, torch.Tensor):
            assert isinstance(b, torch.Tensor) and a.shape == b.shape and a.dtype == b.dtype, path
            if a.is_floating_point():
                torch.testing.assert_close(a, b, rtol=1e-6, atol=1e-7, msg=lambda m: f"{path}: {m}")   # thread-count noise only
            else:
                assert torch.equal(a, b), path
        elif isinstance(a, dict):
            assert set(a) == set(b), path
            for k in a:
                same(a[k], b[k], f"{path}/{k}")
        elif isinstance(a, (list, tuple)):
            assert len(a) == len(b), path
            for i, (x, y) in enumerate(zip(a, b)):
                same(x, y, f"{path}[{i}]")
        else:
            assert a == b, path

    gold = os.path.join(root, "tests", "golden")
    names = sorted(f for f in os.listdir(gold) if f.endswith(".pt"))
    assert names == sorted(f for f in os.listdir(tmp_path) if f.endswith(".pt"))
    for f in names:
        same(torch.load(os.path.join(gold, f), weights_only=False), torch.load(os.path.join(tmp_path, f), weights_only=False), f)
