"""nn.Module mirrors of the reference's hot-path modules, backed by the B200 engine.

``Mamba`` mirrors /root/reference/vim-mamba_ssm/mamba_ssm/modules/mamba_simple.py:34-399 — identical constructor
signature, parameter names/shapes (state-dict compatible), init distributions and ``forward(hidden_states,
inference_params=None)`` contract — and ``RMSNorm`` mirrors vim-mamba_ssm/mamba_ssm/ops/triton/layernorm.py:481-502.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import _lib as L
from . import mixer, ops


class Mamba(nn.Module):
    def __init__(self, d_model, d_state=16, d_conv=4, expand=2, dt_rank="auto", dt_min=0.001, dt_max=0.1,
                 dt_init="random", dt_scale=1.0, dt_init_floor=1e-4, conv_bias=True, bias=False,
                 use_fast_path=True, layer_idx=None, device=None, dtype=None, bimamba_type="none",
                 if_devide_out=False, init_layer_scale=None):
        super().__init__()
        fk = {"device": device, "dtype": dtype}
        self.d_model, self.d_state, self.d_conv, self.expand = d_model, d_state, d_conv, expand
        self.d_inner = int(expand * d_model)
        self.dt_rank = math.ceil(d_model / 16) if dt_rank == "auto" else dt_rank
        self.use_fast_path, self.layer_idx = use_fast_path, layer_idx
        self.bimamba_type, self.if_devide_out = bimamba_type, if_devide_out
        self.init_layer_scale = init_layer_scale
        if bimamba_type not in ("none", "v1", "v2"):
            raise ValueError(f"bimamba_type must be 'none', 'v1' or 'v2', got {bimamba_type!r}")
        if init_layer_scale is not None:
            self.gamma = nn.Parameter(init_layer_scale * torch.ones(d_model))
        self.in_proj = nn.Linear(d_model, 2 * self.d_inner, bias=bias, **fk)

        def branch(suffix):
            conv = nn.Conv1d(self.d_inner, self.d_inner, d_conv, groups=self.d_inner, padding=d_conv - 1,
                             bias=conv_bias, **fk)
            xp = nn.Linear(self.d_inner, self.dt_rank + 2 * d_state, bias=False, **fk)
            dtp = nn.Linear(self.dt_rank, self.d_inner, bias=True, **fk)
            std = self.dt_rank ** -0.5 * dt_scale
            if dt_init == "constant":
                nn.init.constant_(dtp.weight, std)
            elif dt_init == "random":
                nn.init.uniform_(dtp.weight, -std, std)
            else:
                raise NotImplementedError(dt_init)
            # softplus(bias) log-uniform in [dt_min, dt_max]  (reference :103-113)
            t = torch.exp(torch.rand(self.d_inner, **fk) * (math.log(dt_max) - math.log(dt_min)) + math.log(dt_min))
            t = t.clamp(min=dt_init_floor)
            with torch.no_grad():
                dtp.bias.copy_(t + torch.log(-torch.expm1(-t)))
            dtp.bias._no_reinit = True
            setattr(self, "conv1d" + suffix, conv)
            setattr(self, "x_proj" + suffix, xp)
            setattr(self, "dt_proj" + suffix, dtp)

        def s4d_real():
            a = torch.arange(1, d_state + 1, dtype=torch.float32, device=device).repeat(self.d_inner, 1)
            p = nn.Parameter(torch.log(a).contiguous())
            p._no_weight_decay = True
            return p

        def skip():
            p = nn.Parameter(torch.ones(self.d_inner, device=device))
            p._no_weight_decay = True
            return p

        self.activation = "silu"
        self.act = nn.SiLU()
        branch("")
        self.A_log = s4d_real()
        self.D = skip()
        if bimamba_type in ("v1", "v2"):
            self.A_b_log = s4d_real()
        if bimamba_type == "v2":
            branch("_b")
            self.D_b = skip()
        self.out_proj = nn.Linear(self.d_inner, d_model, bias=bias, **fk)

    def forward(self, hidden_states, inference_params=None):
        """hidden_states: (B, L, D) -> (B, L, D)   (reference :169-173)."""
        if inference_params is not None:
            raise NotImplementedError("autoregressive decode (inference_params) is outside the AuM hot path")
        if not self.use_fast_path:
            raise NotImplementedError("use_fast_path=False has no B200 implementation (it ignores bimamba_type upstream)")
        if torch.is_grad_enabled() and (hidden_states.requires_grad or any(p.requires_grad for p in self.parameters())):
            from .autograd import mamba_mixer_autograd
            return mamba_mixer_autograd(self, hidden_states)
        x = hidden_states
        if torch.is_autocast_enabled():
            x = x.to(torch.get_autocast_dtype("cuda"))
        return mixer.mamba_mixer_forward(self, x)

    def step(self, hidden_states, conv_state, ssm_state):
        raise NotImplementedError("single-token decode is outside the AuM hot path")

    def allocate_inference_cache(self, batch_size, max_seqlen, dtype=None, **kwargs):
        dev = self.out_proj.weight.device
        cdt = self.conv1d.weight.dtype if dtype is None else dtype
        sdt = self.dt_proj.weight.dtype if dtype is None else dtype
        return (torch.zeros(batch_size, self.d_model * self.expand, self.d_conv, device=dev, dtype=cdt),
                torch.zeros(batch_size, self.d_model * self.expand, self.d_state, device=dev, dtype=sdt))


def rms_norm_fn(x, weight, bias, residual=None, prenorm=False, residual_in_fp32=False, eps=1e-6):
    """Fused add + RMSNorm (reference layernorm.py:477).  Forward only on this path (see autograd module)."""
    if torch.is_grad_enabled() and (x.requires_grad or weight.requires_grad or (residual is not None and residual.requires_grad)):
        from .autograd import rms_norm_autograd
        return rms_norm_autograd(x, weight, bias, residual, prenorm, residual_in_fp32, eps)
    shape = x.shape
    x2 = x.reshape(-1, shape[-1])
    if x2.stride(-1) != 1:
        x2 = x2.contiguous()
    r2 = None
    if residual is not None:
        r2 = residual.reshape(-1, shape[-1])
        if r2.stride(-1) != 1:
            r2 = r2.contiguous()
    rdt = r2.dtype if r2 is not None else (torch.float32 if residual_in_fp32 else x.dtype)
    res = ops.add_rmsnorm(x2, mixer._f32(weight), mixer._f32(bias) if bias is not None else None, r2, eps=eps,
                          prenorm=prenorm, residual_dtype=rdt)
    if prenorm:
        return res[0].view(shape), res[1].view(shape)
    return res.view(shape)


def layer_norm_fn(x, weight, bias, residual=None, eps=1e-6, prenorm=False, residual_in_fp32=False, is_rms_norm=False):
    if not is_rms_norm:
        raise NotImplementedError("LayerNorm (non-RMS) is not on the AuM path (AudioMamba uses rms_norm=True)")
    return rms_norm_fn(x, weight, bias, residual, prenorm, residual_in_fp32, eps)


class RMSNorm(nn.Module):
    def __init__(self, hidden_size, eps=1e-5, device=None, dtype=None):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(hidden_size, device=device, dtype=dtype))
        self.register_parameter("bias", None)

    def reset_parameters(self):
        nn.init.ones_(self.weight)

    def forward(self, x, residual=None, prenorm=False, residual_in_fp32=False):
        return rms_norm_fn(x, self.weight, self.bias, residual=residual, eps=self.eps, prenorm=prenorm,
                           residual_in_fp32=residual_in_fp32)
