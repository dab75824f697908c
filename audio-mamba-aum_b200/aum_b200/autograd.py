"""Training path of the hot path: autograd.Function wrappers whose backward runs on the B200 engine.

Reference being replaced: BiMambaInnerFn / MambaInnerFn / MambaInnerFnNoOutProj .backward
(/root/reference/vim-mamba_ssm/mamba_ssm/ops/selective_scan_interface.py:519-603, 367-434, 226-289) and the
Triton layer-norm backward (vim-mamba_ssm/mamba_ssm/ops/triton/layernorm.py:196-290).

Recompute policy = the reference's checkpoint_lvl=1: conv1d_out (u) and delta are recomputed in backward
(:531-534); saved are xz, the x_proj outputs (dt, B|C) and the pre-gate scan output.  Native kernels:
aum_selective_scan_bwd (both directions, one launch), aum_causal_conv1d_bwd, tcgen05 GEMMs for the
activation-gradient products.  The weight-gradient GEMMs (dW = X^T dY: reductions over the token axis of two
token-major operands) are plain library GEMMs and go through torch.matmul (cuBLAS) in this round.
d/dz is the mathematically correct gradient, not the shipped kernel's.
"""
from __future__ import annotations

import torch

from . import _lib as L
from . import mixer, ops


def _t(w: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """W^T as a contiguous `dtype` matrix (cached per parameter version) — the K-contiguous operand of dX = dY @ W."""
    return mixer._cache.get(w, f"wT:{dtype}", lambda p: p.t().to(dtype).contiguous())


def _branch_fwd(xz, Di, N, conv, xproj, dtproj, reverse, act):
    """conv -> x_proj -> dt_proj of one parameter set; returns (u, delta, dt, bc)."""
    B, Lq, _ = xz.shape
    M = B * Lq
    R = dtproj.weight.shape[1]
    Rpad = mixer._round_up(R, 8)
    u = ops.causal_conv1d(xz[..., :Di], mixer._conv_w(conv.weight), mixer._f32(conv.bias) if conv.bias is not None else None,
                          silu=True, reverse=reverse)
    dt = torch.empty((M, Rpad), device=xz.device, dtype=act)
    bc = torch.empty((M, 2 * N), device=xz.device, dtype=torch.float32)
    ops.gemm_tn(u.view(M, Di), mixer._w(xproj.weight, act), out=dt, out2=bc, split=R)
    delta = ops.gemm_tn(dt, mixer._w(dtproj.weight, act, pad_cols=Rpad), k=R, bias=mixer._f32(dtproj.bias),
                        act=L.ACT_SOFTPLUS, out_dtype=torch.float32)
    return u, delta.view(B, Lq, Di), dt, bc.view(B, Lq, 2 * N)


class MambaMixerFn(torch.autograd.Function):
    """One Mamba mixer (Fo-Fo 'none', Fo-Bi 'v1', Bi-Bi 'v2') with a native backward.

    forward(ctx, hidden, module, *params): `params` are the module's parameters in `_param_list(module)` order (they
    are passed explicitly so autograd routes their gradients)."""

    @staticmethod
    def forward(ctx, hidden, m, *params):
        act = hidden.dtype
        B, Lq, Dm = hidden.shape
        M = B * Lq
        Di, N = m.d_inner, m.d_state
        if N != 16:
            raise NotImplementedError("training path supports d_state == 16 (the AuM configuration)")
        h2 = hidden.reshape(M, Dm).contiguous()
        in_b = mixer._f32(m.in_proj.bias) if m.in_proj.bias is not None else None
        xz = ops.gemm_tn(h2, mixer._w(m.in_proj.weight, act), bias=in_b).view(B, Lq, 2 * Di)
        z = xz[..., Di:]
        bt = m.bimamba_type
        u, delta, dt, bc = _branch_fwd(xz, Di, N, m.conv1d, m.x_proj, m.dt_proj, False, act)
        # the forward scan leaves its state checkpoints (every 8 steps) for the backward kernel
        ck_f = ops.scan_bwd_workspace(B, Lq, Di, hidden.device)
        ck_b = ops.scan_bwd_workspace(B, Lq, Di, hidden.device) if bt != "none" else ck_f
        fwd = ops.ScanDirection(u, delta, mixer._neg_exp(m.A_log), bc[..., :N], bc[..., N:], mixer._f32(m.D), ckpt=ck_f)
        bwd, dt_b, bc_b, scale = None, None, None, 1.0
        if bt == "v1":
            bwd = ops.ScanDirection(u, delta, mixer._neg_exp(m.A_b_log), bc[..., :N], bc[..., N:], mixer._f32(m.D), ckpt=ck_b)
        elif bt == "v2":
            ub, deltab, dt_b, bc_b = _branch_fwd(xz, Di, N, m.conv1d_b, m.x_proj_b, m.dt_proj_b, True, act)
            bwd = ops.ScanDirection(ub, deltab, mixer._neg_exp(m.A_b_log), bc_b[..., :N], bc_b[..., N:], mixer._f32(m.D_b),
                                    ckpt=ck_b)
            scale = 0.5 if m.if_devide_out else 1.0
        out_z = torch.empty((B, Lq, Di), device=hidden.device, dtype=act)
        y_pre = torch.empty_like(out_z)
        ops.selective_scan(fwd, bwd, z, out=out_z, out_scale=scale, y_pre=y_pre)
        out_b = mixer._f32(m.out_proj.bias) if m.out_proj.bias is not None else None
        out = ops.gemm_tn(out_z.view(M, Di), mixer._w(m.out_proj.weight, act), bias=out_b).view(B, Lq, Dm)
        ctx.m, ctx.scale = m, scale
        ctx.save_for_backward(h2, xz, dt, bc, y_pre, dt_b if dt_b is not None else dt, bc_b if bc_b is not None else bc,
                              ck_f, ck_b)
        return out

    @staticmethod
    def backward(ctx, dout):
        m, scale = ctx.m, ctx.scale
        h2, xz, dt, bc, y_pre, dt_b, bc_b, ck_f, ck_b = ctx.saved_tensors
        act = xz.dtype
        B, Lq, two_di = xz.shape
        Di, N, M = two_di // 2, m.d_state, B * Lq
        Dm = h2.shape[1]
        dev = xz.device
        bt = m.bimamba_type
        dout2 = dout.reshape(M, Dm).to(act).contiguous()
        z = xz[..., Di:]
        f32 = dict(device=dev, dtype=torch.float32)

        # ---- recompute u, delta (checkpoint_lvl = 1, reference :531-534)
        def recompute(conv, dtproj, dt_, reverse):
            R = dtproj.weight.shape[1]
            u_ = ops.causal_conv1d(xz[..., :Di], mixer._conv_w(conv.weight),
                                   mixer._f32(conv.bias) if conv.bias is not None else None, silu=True, reverse=reverse)
            delta_ = ops.gemm_tn(dt_, mixer._w(dtproj.weight, act, pad_cols=dt_.shape[1]), k=R,
                                 bias=mixer._f32(dtproj.bias), act=L.ACT_SOFTPLUS, out_dtype=torch.float32)
            return u_, delta_.view(B, Lq, Di)

        u, delta = recompute(m.conv1d, m.dt_proj, dt, False)

        # ---- out_proj backward: d(out_z) = dout @ W_out            (reference :539-540)
        dout_z = ops.gemm_tn(dout2, _t(m.out_proj.weight, act)).view(B, Lq, Di)

        # ---- scan backward, both directions in one launch           (reference :541-561)
        dxz = torch.empty_like(xz)                      # dx | dz written in place (:537-538)
        dz = dxz[..., Di:]
        out_z = torch.empty((B, Lq, Di), device=dev, dtype=act)
        A = mixer._neg_exp(m.A_log)
        dA = torch.zeros((Di, N), **f32)
        dD = torch.zeros((Di,), **f32)
        du = torch.empty((B, Lq, Di), **f32)
        ddelta = torch.empty((B, Lq, Di), **f32)
        dbc = torch.zeros((B, Lq, 2 * N), **f32)
        d_f = ops.ScanBwdDirection(u, delta, A, bc, mixer._f32(m.D), du, ddelta, dA, dD, dbc, ck_f, ckpt_valid=True)
        d_b = None
        grads_b = {}
        if bt == "v1":
            dA_b = torch.zeros((Di, N), **f32)
            d_b = ops.ScanBwdDirection(u, delta, mixer._neg_exp(m.A_b_log), bc, mixer._f32(m.D), du, ddelta, dA_b, dD, dbc,
                                       ck_b, ckpt_valid=True)
        elif bt == "v2":
            ub, deltab = recompute(m.conv1d_b, m.dt_proj_b, dt_b, True)
            dA_b = torch.zeros((Di, N), **f32)
            dD_b = torch.zeros((Di,), **f32)
            du_b = torch.empty((B, Lq, Di), **f32)
            ddelta_b = torch.empty((B, Lq, Di), **f32)
            dbc_b = torch.zeros((B, Lq, 2 * N), **f32)
            d_b = ops.ScanBwdDirection(ub, deltab, mixer._neg_exp(m.A_b_log), bc_b, mixer._f32(m.D_b), du_b, ddelta_b,
                                       dA_b, dD_b, dbc_b, ck_b, ckpt_valid=True)
        # ddelta comes back already multiplied by softplus'(pre) = 1 - exp(-delta): it IS d(pre-activation)
        ops.selective_scan_bwd(d_f, d_b, z, y_pre, dout_z, dz, out_z, out_scale=scale, softplus_grad=True)

        g = {}
        # ---- out_proj weight grads (library GEMM: reduction over tokens)       (reference :563-564)
        g["out_proj.weight"] = torch.matmul(dout2.t(), out_z.view(M, Di)).float()
        if m.out_proj.bias is not None:
            g["out_proj.bias"] = dout2.float().sum(0)

        # ---- per-branch chain: delta -> dt_proj -> x_proj -> conv                (reference :566-596)
        dxc = torch.empty((B, Lq, Di), device=dev, dtype=act)      # dx of the causal branch
        dxc_b = torch.empty_like(dxc) if bt == "v2" else None

        def branch_bwd(sfx, conv, xproj, dtproj, u_, delta_, dt_, du_, ddelta_, dbc_, reverse, dx_out):
            R = dtproj.weight.shape[1]
            Rpad = dt_.shape[1]
            dpre = ddelta_.view(M, Di)                      # gradient w.r.t. dt_proj's pre-activation (see above)
            g[f"dt_proj{sfx}.bias"] = dpre.sum(0)
            dpre_h = dpre.to(act)
            g[f"dt_proj{sfx}.weight"] = torch.matmul(dpre_h.t(), dt_[:, :R]).float()          # (Di, R)
            # dx_dbl = [d(dt) | dB | dC]   (M, R+2N)
            dxdbl = torch.empty((M, mixer._round_up(R + 2 * N, 8)), device=dev, dtype=act)
            ops.gemm_tn(dpre_h, _t(dtproj.weight, act), out=dxdbl[:, :R])                     # d(dt) = dpre @ W_dt
            dxdbl[:, R:R + 2 * N] = dbc_.view(M, 2 * N).to(act)
            if dxdbl.shape[1] > R + 2 * N:
                dxdbl[:, R + 2 * N:] = 0
            g[f"x_proj{sfx}.weight"] = torch.matmul(dxdbl[:, :R + 2 * N].t(), u_.view(M, Di)).float()   # (R+2N, Di)
            # d(conv_out) = du_scan + dx_dbl @ W_x
            wxT = mixer._cache.get(xproj.weight, f"wT_pad:{act}:{dxdbl.shape[1]}",
                                   lambda p: torch.nn.functional.pad(p.t().to(act), (0, dxdbl.shape[1] - p.shape[0])).contiguous())
            du_x = ops.gemm_tn(dxdbl, wxT, out_dtype=torch.float32)                              # (M, Di) fp32
            dw = torch.zeros((Di, conv.weight.shape[-1]), **f32)
            db = torch.zeros((Di,), **f32) if conv.bias is not None else None
            ops.causal_conv1d_bwd(xz[..., :Di], mixer._conv_w(conv.weight),
                                  mixer._f32(conv.bias) if conv.bias is not None else None,
                                  du_.view(B, Lq, Di), dx_out, dw, db, silu=True, reverse=reverse,
                                  dout2=du_x.view(B, Lq, Di))                       # sums both terms on the fly
            g[f"conv1d{sfx}.weight"] = dw.view(conv.weight.shape)
            if db is not None:
                g[f"conv1d{sfx}.bias"] = db

        branch_bwd("", m.conv1d, m.x_proj, m.dt_proj, u, delta, dt, du, ddelta, dbc, False, dxz[..., :Di] if bt != "v2" else dxc)
        g["A_log"] = dA * A                               # A = -exp(A_log)  =>  dA_log = dA * A
        g["D"] = dD
        if bt == "v1":
            g["A_b_log"] = dA_b * mixer._neg_exp(m.A_b_log)
        elif bt == "v2":
            branch_bwd("_b", m.conv1d_b, m.x_proj_b, m.dt_proj_b, ub, deltab, dt_b, du_b, ddelta_b, dbc_b, True, dxc_b)
            g["A_b_log"] = dA_b * mixer._neg_exp(m.A_b_log)
            g["D_b"] = dD_b
            torch.add(dxc, dxc_b, out=dxz[..., :Di])

        # ---- in_proj backward                                                    (reference mamba_simple.py:185-191)
        dxz2 = dxz.view(M, 2 * Di)
        dhidden = ops.gemm_tn(dxz2, _t(m.in_proj.weight, act)).view(B, Lq, Dm)
        g["in_proj.weight"] = torch.matmul(dxz2.t(), h2).float()
        if m.in_proj.bias is not None:
            g["in_proj.bias"] = dxz2.float().sum(0)

        names = [n for n, _ in _param_list(m)]
        return (dhidden, None) + tuple(g.get(n) for n in names)


def _param_list(m):
    """(name, parameter) pairs of a Mamba module in a fixed order (gamma is applied outside this Function)."""
    return [(n, p) for n, p in m.named_parameters() if n != "gamma"]


def mamba_mixer_autograd(module, hidden_states):
    x = hidden_states
    if torch.is_autocast_enabled():
        x = x.to(torch.get_autocast_dtype("cuda"))
    params = [p for _, p in _param_list(module)]
    out = MambaMixerFn.apply(x, module, *params)
    gamma = getattr(module, "gamma", None)
    if getattr(module, "init_layer_scale", None) is not None and gamma is not None:
        out = out * gamma.to(out.dtype)
    return out


class AddRMSNormFn(torch.autograd.Function):
    """Fused add + RMSNorm, forward and backward on the engine.
    Semantics of LayerNormFn (layernorm.py:380-461) with is_rms_norm=True."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, prenorm, residual_in_fp32, eps):
        shape = x.shape
        x2 = x.reshape(-1, shape[-1]).contiguous()
        r2 = residual.reshape(-1, shape[-1]).contiguous() if residual is not None else None
        rdt = r2.dtype if r2 is not None else (torch.float32 if residual_in_fp32 else x.dtype)
        rstd = torch.empty(x2.shape[0], device=x.device, dtype=torch.float32)
        y, res = ops.add_rmsnorm(x2, mixer._f32(weight), mixer._f32(bias) if bias is not None else None, r2, eps=eps,
                                 prenorm=True, residual_dtype=rdt, rstd=rstd)
        ctx.save_for_backward(res, weight, rstd)
        ctx.has_res, ctx.has_bias, ctx.prenorm = residual is not None, bias is not None, prenorm
        ctx.x_dtype, ctx.r_dtype = x.dtype, (residual.dtype if residual is not None else None)
        y = y.view(shape)
        return (y, res.view(shape)) if prenorm else y

    @staticmethod
    def backward(ctx, dy, *args):
        res, weight, rstd = ctx.saved_tensors
        dim = res.shape[-1]
        dro = args[0] if (ctx.prenorm and args and args[0] is not None) else None
        native = (not ctx.has_bias and res.dtype == torch.float32 and dim % 8 == 0 and dim <= 2048
                  and (dro is None or dro.dtype == torch.float32) and ctx.r_dtype in (None, torch.float32))
        if native:      # CUDA kernel (aum_add_rmsnorm_bwd)
            dy2 = dy.reshape(-1, dim)
            if dy2.dtype != ctx.x_dtype:
                dy2 = dy2.to(ctx.x_dtype)
            dy2 = dy2.contiguous()
            dro2 = dro.reshape(-1, dim).contiguous() if dro is not None else None
            dw = torch.zeros(dim, device=dy.device, dtype=torch.float32)
            dx, dri = ops.add_rmsnorm_bwd(dy2, dro2, res, rstd, mixer._f32(weight), dw, want_dres_in=ctx.has_res)
            return (dx.view(dy.shape), dw.to(weight.dtype), None, dri.view(dy.shape) if dri is not None else None,
                    None, None, None)
        # generic shapes / dtypes: same formula with torch ops on the GPU
        dyf = dy.reshape(-1, dim).float()
        r = res.float()
        xhat = r * rstd[:, None]
        wdy = dyf * weight.float()
        c1 = (xhat * wdy).mean(dim=-1, keepdim=True)
        dr = (wdy - xhat * c1) * rstd[:, None]
        if dro is not None:
            dr = dr + dro.reshape(-1, dim).float()
        dw = (dyf * xhat).sum(0).to(weight.dtype)
        db = dyf.sum(0) if ctx.has_bias else None
        dx = dr.to(ctx.x_dtype).view(dy.shape)
        dres = dr.to(ctx.r_dtype).view(dy.shape) if ctx.has_res else None
        return dx, dw, db, dres, None, None, None


def rms_norm_autograd(x, weight, bias, residual, prenorm, residual_in_fp32, eps):
    return AddRMSNormFn.apply(x, weight, bias, residual, prenorm, residual_in_fp32, eps)
