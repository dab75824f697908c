"""The C-ABI library loads on a CPU-only box and exports every symbol include/aum_b200.h declares
(no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "aum_b200.h")


def _declared():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"AUM_API\s+[\w\s\*]+?\b(aum_\w+)\s*\(", src)))


def test_header_declares_the_hot_path_entry_points():
    names = _declared()
    for must in ("aum_gemm_tn", "aum_causal_conv1d_fwd", "aum_selective_scan_fwd", "aum_add_rmsnorm_fwd",
                 "aum_transpose", "aum_last_error", "aum_version"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from aum_b200 import _lib
    if not os.path.isfile(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(L, name), f"{name} declared in aum_b200.h but not exported"
    # the python binding table covers exactly the declared API
    assert sorted(_lib.SIGNATURES) == _declared()
    lib = _lib.lib()
    assert lib.aum_version() == 100


def test_argument_errors_are_reported_not_crashed():
    from aum_b200 import _lib
    lib = _lib.lib()
    rc = lib.aum_causal_conv1d_fwd(None, 0, None, None, None, 0, 1, 1, 1, 4, 0, 1, 0, None)
    assert rc != 0 and b"null" in lib.aum_last_error()
    rc = lib.aum_selective_scan_fwd(None, None, None, 0, None, 0, 1, 1, 1, 16, 0, 1.0, None, 0, 0, None)
    assert rc != 0 and b"direction" in lib.aum_last_error()


def test_product_path_has_no_cpu_fallback():
    import torch
    from aum_b200 import ops, AumError
    with pytest.raises(AumError):
        ops.causal_conv1d(torch.zeros(1, 4, 8), torch.zeros(8, 4), None)


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "audio-mamba-aum_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(dp, fn)).read()
                assert "aum_oracle" not in txt and "ref_loader" not in txt, os.path.join(dp, fn)


def test_mamba_module_state_dict_keys_match_reference_golden():
    import torch
    from mamba_ssm.modules.mamba_simple import Mamba
    from conftest import load_golden
    for name, c in load_golden("mamba_module.pt").items():
        m = Mamba(c["d_model"], **c["kwargs"])
        ours = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        ref = {k: tuple(v.shape) for k, v in c["state"].items()}
        assert ours == ref, name
        assert list(m.state_dict().keys()) == list(c["state"].keys()), name
        m.load_state_dict(c["state"], strict=True)
    m = Mamba(64, bimamba_type="v1")
    assert m.A_log._no_weight_decay and m.D._no_weight_decay and m.dt_proj.bias._no_reinit
    assert torch.allclose(m.A_log[0], torch.log(torch.arange(1, 17.0)))
