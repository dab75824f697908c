"""The reference's functional operator API on the B200 engine — same names, argument order and
(batch, channel, length) layouts as /root/reference/vim-mamba_ssm/mamba_ssm/ops/selective_scan_interface.py:
selective_scan_fn (:77-83), mamba_inner_fn (:606-614), bimamba_inner_fn (:616-624),
mamba_inner_fn_no_out_proj (:627-633) and causal_conv1d_fn (pip causal_conv1d, used at :646,:683).

The engine is token-major.  A (B, C, L) argument that is really a transposed view of a token-major buffer
(stride(1) == 1 — what this package's own Mamba module passes) is used in place; a genuinely channel-major
tensor (stride(-1) == 1, the reference's layout, :458-459) goes through one aum_transpose launch.
Like the reference's autograd.Functions (:14-74, 155-289, 292-434, 437-603) every op here is differentiable: when a
gradient is required the call is routed through the Functions of aum_b200.autograd (native backward kernels), with
the layout adapters done by differentiable torch views / copies.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib as L
from . import mixer, ops


def _needs_grad(*ts) -> bool:
    return torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in ts)


def _tm_grad(t: torch.Tensor, dtype=None) -> torch.Tensor:
    """(B, C, L) -> token-major (B, L, C) with differentiable torch ops (autograd path)."""
    tm = t.transpose(1, 2)
    if dtype is not None and tm.dtype != dtype:
        tm = tm.to(dtype)
    if tm.stride(-1) != 1 or (tm.shape[0] > 1 and tm.stride(0) != tm.shape[1] * tm.stride(1)):
        tm = tm.contiguous()
    return tm


def _to_token_major(t: torch.Tensor, dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """(B, C, L) -> (B, L, C) token-major tensor of `dtype` (view when possible)."""
    dtype = dtype or t.dtype
    B, Cc, Lq = t.shape
    if t.dtype == dtype and (t.stride(1) == 1 or Cc == 1) and (B == 1 or t.stride(0) == Lq * t.stride(2)) and t.stride(2) >= Cc:
        return t.transpose(1, 2)
    if t.stride(2) != 1 and Lq != 1:
        t = t.contiguous()
    return ops.transpose(t, dst_dtype=dtype)


def _to_channel_major(t_tm: torch.Tensor, as_view: bool) -> torch.Tensor:
    """(B, L, C) token-major -> (B, C, L)."""
    return t_tm.transpose(1, 2) if as_view else ops.transpose(t_tm)


def _is_tm_view(t: torch.Tensor) -> bool:
    return t.stride(1) == 1 and t.stride(2) != 1


def causal_conv1d_fn(x, weight, bias=None, activation=None):
    """x: (B, D, L); weight: (D, W); bias: (D,); activation in (None, 'silu', 'swish').  Returns (B, D, L)."""
    if activation not in (None, "silu", "swish"):
        raise NotImplementedError("activation must be None, silu or swish")
    if _needs_grad(x, weight, bias):
        from .autograd import ConvFn
        return ConvFn.apply(_tm_grad(x), weight, bias, activation is not None).transpose(1, 2)
    view = _is_tm_view(x)
    x_tm = _to_token_major(x)
    y = ops.causal_conv1d(x_tm, mixer._conv_w(weight), mixer._f32(bias) if bias is not None else None,
                          silu=activation is not None)
    return _to_channel_major(y, view)


def selective_scan_fn(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                      return_last_state=False):
    """u, delta, z: (B, D, L); A: (D, N) real; B, C: (B, N, L) or (B, 1, N, L); D, delta_bias: (D,).
    Returns out (B, D, L) [and last_state (B, D, N)]  (reference :77-83, semantics :86-152)."""
    if A.is_complex():
        raise NotImplementedError("complex A is never used by AuM (mamba_simple.py:193)")
    if B.dim() == 4:
        if B.shape[1] != 1:
            raise NotImplementedError("grouped B/C (G>1) is not used by AuM")
        B = B[:, 0]
    if C.dim() == 4:
        if C.shape[1] != 1:
            raise NotImplementedError("grouped B/C (G>1) is not used by AuM")
        C = C[:, 0]
    if B.dim() != 3 or C.dim() != 3:
        raise NotImplementedError("only input-dependent B/C of shape (B, N, L) are supported (the AuM mode)")
    if _needs_grad(u, delta, A, B, C, D, z, delta_bias):
        return _selective_scan_grad(u, delta, A, B, C, D, z, delta_bias, delta_softplus, return_last_state)
    view = _is_tm_view(u)
    dt_ = u.dtype
    u_tm = _to_token_major(u)
    delta_tm = _to_token_major(delta if delta.dtype in (dt_, torch.float32) else delta.to(dt_))
    z_tm = _to_token_major(z, dt_) if z is not None else None
    B_tm = _to_token_major(B)
    C_tm = _to_token_major(C, B_tm.dtype)
    Bsz, Lq, Dch = u_tm.shape
    N = A.shape[1]
    last = torch.empty((Bsz, Dch, N), device=u.device, dtype=torch.float32) if return_last_state else None
    d = ops.ScanDirection(u_tm, delta_tm, mixer._f32(A), B_tm, C_tm,
                          mixer._f32(D) if D is not None else None,
                          mixer._f32(delta_bias) if delta_bias is not None else None,
                          delta_softplus, last)
    out_tm = ops.selective_scan(d, None, z_tm)
    out = _to_channel_major(out_tm, view)
    return (out, last) if return_last_state else out


def _selective_scan_grad(u, delta, A, B, C, D, z, delta_bias, delta_softplus, return_last_state):
    """selective_scan_fn under autograd (SelectiveScanFn, :14-74): bias + softplus and the layout adapters as
    differentiable torch ops, the recurrence and its backward on the engine (autograd.ScanFn).  d_state == 16."""
    from .autograd import ScanFn
    if A.shape[1] != 16:
        raise NotImplementedError("the differentiable scan supports d_state == 16 (the AuM configuration)")
    dt_ = u.dtype
    d32 = delta.float()
    if delta_bias is not None:
        d32 = d32 + delta_bias.float()[None, :, None]
    if delta_softplus:
        d32 = torch.nn.functional.softplus(d32)
    bc = torch.cat([B.float().transpose(1, 2), C.float().transpose(1, 2)], dim=-1).contiguous()      # (B, L, 2N)
    out_tm = ScanFn.apply(_tm_grad(u), _tm_grad(d32), A.float().contiguous(), bc,
                          D.float().contiguous() if D is not None else None, _tm_grad(z, dt_) if z is not None else None)
    out = out_tm.transpose(1, 2)
    if not return_last_state:
        return out
    with torch.no_grad():        # the reference returns last_state without a gradient path as well (:40-42)
        _, last = selective_scan_fn(u.detach(), delta.detach(), A.detach(), B.detach(), C.detach(),
                                    D.detach() if D is not None else None, z.detach() if z is not None else None,
                                    delta_bias.detach() if delta_bias is not None else None, delta_softplus, True)
    return out, last


def _autocast_dtype(x: torch.Tensor) -> torch.dtype:
    return torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled() else x.dtype


class _P:
    """Light parameter holder so the functional entry points can reuse mixer._pipeline."""
    pass


def _inner(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, A, A_b, D, delta_bias,
           B, C, B_proj_bias, C_proj_bias, delta_softplus, two_dirs: bool):
    if B is not None or C is not None or B_proj_bias is not None or C_proj_bias is not None:
        raise NotImplementedError("only input-dependent B/C without projection biases (the AuM mode, mamba_simple.py:208-209)")
    if not delta_softplus:
        raise NotImplementedError("delta_softplus=False is never used by the inner ops (mamba_simple.py:212)")
    if A.is_complex():
        raise NotImplementedError("complex A is never used by AuM")
    act = _autocast_dtype(xz)
    xz_tm = _to_token_major(xz, act)            # (B, L, 2Di)
    Di = xz_tm.shape[-1] // 2
    N = A.shape[-1]
    u, delta, Bm, Cm = mixer._pipeline(xz_tm, Di, N, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight,
                                       delta_bias if delta_bias is not None else torch.zeros(Di, device=xz.device),
                                       reverse=False, delta_dtype=torch.float32, backend=L.GEMM_AUTO)
    Dv = mixer._f32(D) if D is not None else None
    fwd = ops.ScanDirection(u, delta, mixer._f32(A), Bm, Cm, Dv)
    bwd = ops.ScanDirection(u, delta, mixer._f32(A_b), Bm, Cm, Dv) if two_dirs else None
    return ops.selective_scan(fwd, bwd, xz_tm[..., Di:]), _is_tm_view(xz)


def _inner_grad(mode, has_out, xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, out_proj_weight,
                out_proj_bias, A, A_b, D, delta_bias, B, C, B_proj_bias, C_proj_bias, delta_softplus):
    """The fused inner ops under autograd (BiMambaInnerFn / MambaInnerFn / MambaInnerFnNoOutProj): autograd.InnerFn on a
    token-major xz; gradients w.r.t. xz come back through the transpose as the reference's dxz (:599)."""
    from .autograd import InnerCfg, InnerFn
    if B is not None or C is not None or B_proj_bias is not None or C_proj_bias is not None:
        raise NotImplementedError("only input-dependent B/C without projection biases (the AuM mode, mamba_simple.py:208-209)")
    if not delta_softplus:
        raise NotImplementedError("delta_softplus=False is never used by the inner ops (mamba_simple.py:212)")
    if A.is_complex():
        raise NotImplementedError("complex A is never used by AuM")
    act = _autocast_dtype(xz)
    xz_tm = _tm_grad(xz, act)
    cfg = InnerCfg(mode, has_out, 1.0, False)
    return InnerFn.apply(cfg, xz_tm, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, delta_bias, A, D,
                         None, None, None, None, None, A_b, None, out_proj_weight, out_proj_bias)


def mamba_inner_fn(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight,
                   out_proj_weight, out_proj_bias, A, B=None, C=None, D=None, delta_bias=None,
                   B_proj_bias=None, C_proj_bias=None, delta_softplus=True):
    """Fo-Fo fused inner op (reference :606-614 / MambaInnerFn.forward :296-365).  xz: (B, 2Di, L) -> (B, L, Dm)."""
    if _needs_grad(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, out_proj_weight, out_proj_bias, A, D, delta_bias):
        return _inner_grad("none", True, xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, out_proj_weight,
                           out_proj_bias, A, None, D, delta_bias, B, C, B_proj_bias, C_proj_bias, delta_softplus)
    out_z, _ = _inner(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, A, None, D, delta_bias,
                      B, C, B_proj_bias, C_proj_bias, delta_softplus, two_dirs=False)
    Bsz, Lq, Di = out_z.shape
    ob = mixer._f32(out_proj_bias) if out_proj_bias is not None else None
    return ops.gemm_tn(out_z.view(Bsz * Lq, Di), mixer._w(out_proj_weight, out_z.dtype), bias=ob).view(Bsz, Lq, -1)


def bimamba_inner_fn(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight,
                     out_proj_weight, out_proj_bias, A, A_b, B=None, C=None, D=None, delta_bias=None,
                     B_proj_bias=None, C_proj_bias=None, delta_softplus=True):
    """Fo-Bi fused inner op (reference :616-624 / BiMambaInnerFn.forward :441-517).  xz: (B, 2Di, L) -> (B, L, Dm)."""
    if _needs_grad(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, out_proj_weight, out_proj_bias, A, A_b, D, delta_bias):
        return _inner_grad("v1", True, xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, out_proj_weight,
                           out_proj_bias, A, A_b, D, delta_bias, B, C, B_proj_bias, C_proj_bias, delta_softplus)
    out_z, _ = _inner(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, A, A_b, D, delta_bias,
                      B, C, B_proj_bias, C_proj_bias, delta_softplus, two_dirs=True)
    Bsz, Lq, Di = out_z.shape
    ob = mixer._f32(out_proj_bias) if out_proj_bias is not None else None
    return ops.gemm_tn(out_z.view(Bsz * Lq, Di), mixer._w(out_proj_weight, out_z.dtype), bias=ob).view(Bsz, Lq, -1)


def mamba_inner_fn_no_out_proj(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight,
                               A, B=None, C=None, D=None, delta_bias=None, B_proj_bias=None, C_proj_bias=None,
                               delta_softplus=True):
    """One Bi-Bi pipeline (reference :627-633 / MambaInnerFnNoOutProj.forward :159-224).  Returns (B, Di, L)."""
    if _needs_grad(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, A, D, delta_bias):
        return _inner_grad("none", False, xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, None, None,
                           A, None, D, delta_bias, B, C, B_proj_bias, C_proj_bias, delta_softplus).transpose(1, 2)
    out_z, view = _inner(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, A, None, D, delta_bias,
                         B, C, B_proj_bias, C_proj_bias, delta_softplus, two_dirs=False)
    return _to_channel_major(out_z, view)
