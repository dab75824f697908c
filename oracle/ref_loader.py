"""Import the REAL reference (kaistmm/Audio-Mamba-AuM) on CPU, unmodified.

TEST / BENCH INFRASTRUCTURE ONLY (never imported by the product).  Source tree, first that exists:
``$AUM_REFERENCE_ROOT``, ``/root/reference`` (build container), ``oracle/_ref`` (the byte-for-byte staged copy
of the needed files made by ``oracle/build_ref.py``; git-ignored, travels to the GPU box).  Nothing here reads
/root/reference on the GPU box.  Used by ``oracle/make_golden.py`` (fixture generation), the CPU oracle tests,
``bench.py --impl reference`` / ``cpu_baseline`` (the reference's own CPU path, timed) and the drop-in tests.

Recipe: the reference's hot path imports three native pip modules
unconditionally (selective_scan_interface.py:9-11).  We register empty stand-ins so the *Python*
reference functions (``selective_scan_ref``, ``bimamba_inner_ref``, ``mamba_inner_ref``,
``rms_norm_ref``, ``AudioMamba``) import and run unmodified; then rebind, in the loaded module's
namespace only, the two names the ``*_inner_ref`` bodies call into CUDA through:
``causal_conv1d_fn`` := the reference's own fallback expression (mamba_simple.py:272,:82) and
``selective_scan_fn`` := the reference's own ``selective_scan_ref``.  No reference source is
copied; every function body that produces a fixture is the reference's.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

import torch
import torch.nn.functional as F

_PROBE = "vim-mamba_ssm/mamba_ssm/ops/selective_scan_interface.py"


def _find_root() -> str:
    cands = [os.environ.get("AUM_REFERENCE_ROOT"), "/root/reference",
             os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, _PROBE)):
            return c
    return cands[1]


REF_ROOT = _find_root()
REF_KIND = "staged copy (oracle/_ref)" if REF_ROOT.endswith("_ref") else "reference tree"


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, _PROBE))


def _stub(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def _conv_fallback(x, weight, bias=None, activation=None):
    """The reference's own non-CUDA conv expression: act(conv1d(x)[..., :L]), padding=W-1
    (vim-mamba_ssm/mamba_ssm/modules/mamba_simple.py:272 with the Conv1d of :76-84)."""
    W = weight.shape[-1]
    L = x.shape[-1]
    y = F.conv1d(x, weight[:, None, :], bias, padding=W - 1, groups=x.shape[1])[..., :L]
    return F.silu(y) if activation in ("silu", "swish") else y


_loaded = {}


def load_reference():
    """Returns a namespace with the reference's own python callables/classes."""
    if "ns" in _loaded:
        return _loaded["ns"]
    if not reference_available():
        raise RuntimeError("reference tree not present at " + REF_ROOT)
    import importlib.machinery  # noqa: F401

    # native pip modules the reference imports unconditionally
    _stub("causal_conv1d", causal_conv1d_fn=_conv_fallback, causal_conv1d_update=None)
    _stub("causal_conv1d_cuda")
    _stub("selective_scan_cuda")
    # bare package so sub-modules import without running mamba_ssm/__init__.py
    # (its generation.py needs a transformers symbol that no longer exists)
    pkg_dir = os.path.join(REF_ROOT, "vim-mamba_ssm", "mamba_ssm")
    for name, sub in (("mamba_ssm", ""), ("mamba_ssm.ops", "ops"), ("mamba_ssm.ops.triton", "ops/triton"),
                      ("mamba_ssm.modules", "modules")):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(pkg_dir, sub)]
        m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None, is_package=True)
        sys.modules[name] = m

    ssi = importlib.import_module("mamba_ssm.ops.selective_scan_interface")
    # the *_inner_ref bodies call these two names; point them at the reference's own python refs
    ssi.causal_conv1d_fn = _conv_fallback
    ssi.selective_scan_fn = ssi.selective_scan_ref

    ln = importlib.import_module("mamba_ssm.ops.triton.layernorm")   # triton import only; kernels never run

    def _rms_norm_fn(x, weight, bias, residual=None, prenorm=False, residual_in_fp32=False, eps=1e-6):
        return ln.rms_norm_ref(x, weight, bias, residual=residual, eps=eps, prenorm=prenorm, upcast=True)

    ln.rms_norm_fn = _rms_norm_fn
    ln.RMSNorm.forward = lambda self, x, residual=None, prenorm=False, residual_in_fp32=False: _rms_norm_fn(
        x, self.weight, self.bias, residual=residual, prenorm=prenorm, eps=self.eps)

    ms = importlib.import_module("mamba_ssm.modules.mamba_simple")
    # CPU stand-ins for the CUDA-backed fused ops: the reference's own *_ref functions
    ms.bimamba_inner_fn = ssi.bimamba_inner_ref
    ms.mamba_inner_fn = ssi.mamba_inner_ref

    def _no_out_proj(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, A,
                     B=None, C=None, D=None, delta_bias=None, B_proj_bias=None, C_proj_bias=None,
                     delta_softplus=True):
        # mamba_inner_ref with an identity out_proj (selective_scan_interface.py:636-670) -> (B, Di, L)
        Di = xz.shape[1] // 2
        eye = torch.eye(Di, dtype=xz.dtype)
        y = ssi.mamba_inner_ref(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight,
                                eye, None, A, B, C, D, delta_bias, B_proj_bias, C_proj_bias, delta_softplus)
        return y.transpose(1, 2)

    ms.mamba_inner_fn_no_out_proj = _no_out_proj

    ns = types.SimpleNamespace(ssi=ssi, ln=ln, ms=ms, Mamba=ms.Mamba)
    _loaded["ns"] = ns
    return ns


def _timm_shim():
    """The four timm symbols src/models/mamba_models.py and src/utilities/*.py import (timm is not installed)."""
    import math

    def to_2tuple(x):
        return tuple(x) if isinstance(x, (tuple, list)) else (x, x)

    def trunc_normal_(t, mean=0.0, std=1.0, a=-2.0, b=2.0):
        return torch.nn.init.trunc_normal_(t, mean=mean, std=std, a=a, b=b)

    def lecun_normal_(t):
        fan_in = t.shape[1] * (t[0][0].numel() if t.dim() > 2 else 1)
        return torch.nn.init.trunc_normal_(t, std=math.sqrt(1.0 / fan_in) / .87962566103423978)

    class DropPath(torch.nn.Module):
        def __init__(self, p=0.0):
            super().__init__()
            self.p = p

        def forward(self, x):
            return x

    _stub("timm")
    _stub("timm.models")
    _stub("timm.models.layers", to_2tuple=to_2tuple, trunc_normal_=trunc_normal_,
          lecun_normal_=lecun_normal_, DropPath=DropPath)
    _stub("timm.layers", to_2tuple=to_2tuple, trunc_normal_=trunc_normal_,
          lecun_normal_=lecun_normal_, DropPath=DropPath)
    _stub("wget")


def load_reference_model():
    """Also import src/models/mamba_models.py (AudioMamba) with a 4-symbol timm shim."""
    ns = load_reference()
    if hasattr(ns, "AudioMamba"):
        return ns
    _timm_shim()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    mm = importlib.import_module("src.models.mamba_models")
    mm.rms_norm_fn = ns.ln.rms_norm_fn
    ns.mm = mm
    ns.AudioMamba = mm.AudioMamba
    return ns


def load_reference_model_over_shim():
    """Drop-in check: the reference's OWN ``src/models/mamba_models.py`` (AudioMamba, Block, create_block), unmodified,
    importing ``mamba_ssm.modules.mamba_simple.Mamba`` and ``mamba_ssm.ops.triton.layernorm.{RMSNorm, rms_norm_fn}``
    (mamba_models.py:18,26) from THIS repo's ``audio-mamba-aum_b200/mamba_ssm`` package instead of the pip wheel.
    Must run in a process that has not called load_reference() (both register a package named ``mamba_ssm``)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at " + REF_ROOT)
    if "ns" in _loaded:
        raise RuntimeError("load_reference() already registered the reference's mamba_ssm in this process")
    import importlib.machinery  # noqa: F401
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "audio-mamba-aum_b200")
    if pkg not in sys.path:
        sys.path.insert(0, pkg)
    import mamba_ssm.modules.mamba_simple as shim_ms     # this repo's shim
    assert os.path.abspath(shim_ms.__file__).startswith(pkg), shim_ms.__file__
    _timm_shim()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    mm = importlib.import_module("src.models.mamba_models")
    assert mm.Mamba is shim_ms.Mamba
    return types.SimpleNamespace(mm=mm, AudioMamba=mm.AudioMamba, Mamba=shim_ms.Mamba)
