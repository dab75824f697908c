"""Training path of the hot path: autograd.Functions whose forward AND backward run on the B200 engine.

Reference being replaced: BiMambaInnerFn / MambaInnerFn / MambaInnerFnNoOutProj / SelectiveScanFn (forward + backward,
/root/reference/vim-mamba_ssm/mamba_ssm/ops/selective_scan_interface.py:437-603, 292-434, 155-289, 14-74), the
in_proj matmul under autograd (mamba_simple.py:185-191) and the Triton layer-norm backward
(vim-mamba_ssm/mamba_ssm/ops/triton/layernorm.py:196-290).

Recompute policy = the reference's checkpoint_lvl=1: conv1d_out (u) and delta are recomputed in backward (:531-534);
saved are xz, the x_proj outputs (dt, B|C), the pre-gate scan output and the forward scan's state checkpoints.
Every product is a native kernel: aum_selective_scan_bwd (both directions, one launch), aum_causal_conv1d_bwd,
aum_gemm_tn for the activation gradients (dX = dY W) and aum_gemm_wgrad for the weight gradients (dW = dY^T X: tcgen05
with MN-major operands, fp32 accumulation straight into the gradient buffer).  d/dz is the mathematically correct
gradient, not the shipped kernel's.

Gradient delivery.  By default every Function returns its parameter gradients to autograd (fresh fp32 tensors).  A
trainer that owns a flat gradient buffer (aum_b200.dist.FlatGradReducer) marks its parameters ``_aum_direct_grad``:
the kernels then ACCUMULATE into ``param.grad`` (a view into that buffer) and the Function returns None for them - no
temporary, no cast, no AccumulateGrad pass.
"""
from __future__ import annotations

import os
from collections import namedtuple

import torch

from . import _lib as L
from . import mixer, ops

F32 = torch.float32
_GRAD_16BIT = os.environ.get("AUM_GRAD_16BIT", "1") == "1"


def _t(w: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """W^T as a contiguous `dtype` matrix (cached per parameter version) — the K-contiguous operand of dX = dY @ W."""
    return mixer._wT(w, dtype)


def _direct(p) -> bool:
    """True when this parameter's gradient is accumulated in place into p.grad (see the module docstring)."""
    return (p is not None and getattr(p, "_aum_direct_grad", False) and p.grad is not None
            and p.grad.dtype == F32 and p.grad.is_contiguous())


def _grad_buffer(p, shape2d):
    """(fp32 2-D accumulation target, direct?) for parameter p: its own .grad when direct, else a zeroed temporary."""
    if _direct(p):
        return p.grad.view(shape2d), True
    return torch.zeros(shape2d, device=p.device, dtype=F32), False


def _deliver(p, buf, direct):
    """What the Function returns for p: None when the kernels already accumulated into p.grad."""
    if p is None or direct:
        return None
    return buf.view(p.shape).to(p.dtype)


def _deliver_value(p, val):
    """Small gradients computed as a tensor (bias sums, dA * A): add into p.grad when direct, else return."""
    if p is None:
        return None
    if _direct(p):
        p.grad.add_(val.view(p.shape))
        return None
    return val.view(p.shape).to(p.dtype)


def _wgrad(dy2d, x2d, p):
    """dW = dY^T X for parameter p of shape (No, Ki) (aum_gemm_wgrad)."""
    if p is None or not p.requires_grad:
        return None
    buf, direct = _grad_buffer(p, (dy2d.shape[1], x2d.shape[1]))
    if direct and mixer._WGRAD_SIDE:
        # accumulates into the trainer's flat buffer, which nothing reads before FlatGradReducer joins: second stream
        mixer.side_work.launch(lambda: ops.gemm_wgrad(dy2d, x2d, buf), dy2d, x2d, buf)
        return None
    ops.gemm_wgrad(dy2d, x2d, buf)
    return _deliver(p, buf, direct)


# ------------------------------------------------------------------------------------------------------
class LinearFn(torch.autograd.Function):
    """y[M, N] = x[M, K] @ W[N, K]^T (+ b): the in_proj matmul (mamba_simple.py:185-191) under autograd; also the
    patch-embedding conv as a linear map of im2col rows (weight (N, 1, pf, pt) is used as (N, pf * pt)) and the head.
    out_dtype: dtype of y (default: x's)."""

    @staticmethod
    def forward(ctx, x2, weight, bias, out_dtype=None):
        act = x2.dtype
        w2 = mixer._w2d(weight, act)
        y = ops.gemm_tn(x2, w2, bias=mixer._f32(bias) if bias is not None else None, out_dtype=out_dtype)
        ctx.save_for_backward(x2)
        ctx.weight, ctx.bias = weight, bias
        return y

    @staticmethod
    def backward(ctx, dy):
        (x2,) = ctx.saved_tensors
        w, b = ctx.weight, ctx.bias
        act = x2.dtype
        dy2 = dy if (dy.dtype == act and dy.stride(-1) == 1) else dy.to(act).contiguous()
        dx = ops.gemm_tn(dy2, _t(w, act)) if ctx.needs_input_grad[0] else None
        dw = _wgrad(dy2, x2, w) if ctx.needs_input_grad[1] else None
        db = _deliver_value(b, dy2.float().sum(0)) if (b is not None and ctx.needs_input_grad[2]) else None
        return dx, dw, db, None


# ------------------------------------------------------------------------------------------------------
# mode: 'none' (Fo-Fo, one direction), 'v1' (Fo-Bi: shared projections, A vs A_b), 'v2' (Bi-Bi: two parameter sets,
# the second walking the sequence backwards).  has_out: apply out_proj inside (the *_inner_fn ops) or return out_z
# (mamba_inner_fn_no_out_proj).  a_is_log: the A arguments are A_log parameters (module path: A = -exp(A_log) comes from
# the derived-weight cache and dA_log = dA * A is returned) instead of A itself (functional API).
InnerCfg = namedtuple("InnerCfg", "mode has_out scale a_is_log")


def _A(a, cfg):
    return mixer._neg_exp(a) if cfg.a_is_log else mixer._f32(a)


def _train16(act, Di) -> bool:
    """Training keeps delta (forward and recomputed), du / ddelta of the backward scan and the x_proj term of d(conv_out) in
    the activation dtype when that is 16-bit - the rounding points of the reference's kernels under autocast
    (selective_scan_interface.py:468, :541-561, :590); every sum over those terms is still formed in fp32.  Needs the
    specialised backward-scan kernel (128-channel CTAs); AUM_GRAD_16BIT=0 keeps fp32 everywhere."""
    return (act != F32 and Di % 128 == 0 and _GRAD_16BIT and "AUM_SCAN_BWD_GENERIC" not in os.environ
            and "AUM_SCAN_BWD_NOSPEC" not in os.environ)


def _branch_fwd(xz, Di, N, cw, cb, xw, dtw, dtb, reverse, act):
    """conv -> x_proj -> dt_proj of one parameter set (selective_scan_interface.py:461-496); returns (u, delta, dt, bc)."""
    B, Lq, _ = xz.shape
    M = B * Lq
    R = dtw.shape[1]
    Rpad = mixer._round_up(R, 8)
    dt = torch.empty((M, Rpad), device=xz.device, dtype=act)
    bc = torch.empty((M, 2 * N), device=xz.device, dtype=F32)
    cw_, cb_, wx = mixer._conv_w(cw), (mixer._f32(cb) if cb is not None else None), mixer._w(xw, act)
    if mixer._FUSE_CONV_XPROJ and ops.conv_xproj_eligible(xz[..., :Di], cw_, wx, R, 2 * N):
        u = ops.conv_xproj(xz[..., :Di], cw_, cb_, wx, R, dt, bc, reverse=reverse)
    else:
        u = ops.causal_conv1d(xz[..., :Di], cw_, cb_, silu=True, reverse=reverse)
        ops.gemm_tn(u.view(M, Di), wx, out=dt, out2=bc, split=R)
    delta = ops.gemm_tn(dt, mixer._w(dtw, act, pad_cols=Rpad), k=R,
                        bias=mixer._f32(dtb) if dtb is not None else None, act=L.ACT_SOFTPLUS,
                        out_dtype=act if _train16(act, Di) else F32)
    return u, delta.view(B, Lq, Di), dt, bc.view(B, Lq, 2 * N)


class InnerFn(torch.autograd.Function):
    """conv -> x_proj -> dt_proj -> selective scan(s) [-> out_proj] of one mixer on a token-major xz (B, L, 2 Di).

    forward(ctx, cfg, xz, cw, cb, xw, dtw, dtb, A, D, cw_b, cb_b, xw_b, dtw_b, dtb_b, A_b, D_b, ow, ob)
    (the *_b set only for 'v2', A_b also for 'v1'; absent tensors are None)."""

    @staticmethod
    def forward(ctx, cfg, xz, cw, cb, xw, dtw, dtb, A, D, cw_b, cb_b, xw_b, dtw_b, dtb_b, A_b, D_b, ow, ob):
        act = xz.dtype
        B, Lq, two_di = xz.shape
        Di, M = two_di // 2, B * Lq
        N = A.shape[-1]
        if N != 16:
            raise NotImplementedError("training path supports d_state == 16 (the AuM configuration)")
        if xz.stride(-1) != 1 or (B > 1 and xz.stride(0) != Lq * xz.stride(1)):
            xz = xz.contiguous()
        z = xz[..., Di:]
        u, delta, dt, bc = _branch_fwd(xz, Di, N, cw, cb, xw, dtw, dtb, False, act)
        # the forward scan leaves its state checkpoints (every 8 steps) for the backward kernel
        ck_f = ops.scan_bwd_workspace(B, Lq, Di, xz.device)
        ck_b = ops.scan_bwd_workspace(B, Lq, Di, xz.device) if cfg.mode != "none" else ck_f
        Dv = mixer._f32(D) if D is not None else None
        fwd = ops.ScanDirection(u, delta, _A(A, cfg), bc[..., :N], bc[..., N:], Dv, ckpt=ck_f)
        bwd, dt_b, bc_b = None, None, None
        if cfg.mode == "v1":
            bwd = ops.ScanDirection(u, delta, _A(A_b, cfg), bc[..., :N], bc[..., N:], Dv, ckpt=ck_b)
        elif cfg.mode == "v2":
            ub, deltab, dt_b, bc_b = _branch_fwd(xz, Di, N, cw_b, cb_b, xw_b, dtw_b, dtb_b, True, act)
            bwd = ops.ScanDirection(ub, deltab, _A(A_b, cfg), bc_b[..., :N], bc_b[..., N:],
                                    mixer._f32(D_b) if D_b is not None else None, ckpt=ck_b)
        out_z = torch.empty((B, Lq, Di), device=xz.device, dtype=act)
        y_pre = torch.empty_like(out_z)
        ops.selective_scan(fwd, bwd, z, out=out_z, out_scale=cfg.scale, y_pre=y_pre)
        ctx.cfg = cfg
        ctx.params = (cw, cb, xw, dtw, dtb, A, D, cw_b, cb_b, xw_b, dtw_b, dtb_b, A_b, D_b, ow, ob)
        ctx.save_for_backward(xz, dt, bc, y_pre, dt_b if dt_b is not None else dt, bc_b if bc_b is not None else bc,
                              ck_f, ck_b)
        if not cfg.has_out:
            return out_z
        return ops.gemm_tn(out_z.view(M, Di), mixer._w(ow, act),
                           bias=mixer._f32(ob) if ob is not None else None).view(B, Lq, -1)

    @staticmethod
    def backward(ctx, dout):
        cfg = ctx.cfg
        cw, cb, xw, dtw, dtb, A, D, cw_b, cb_b, xw_b, dtw_b, dtb_b, A_b, D_b, ow, ob = ctx.params
        xz, dt, bc, y_pre, dt_b, bc_b, ck_f, ck_b = ctx.saved_tensors
        act = xz.dtype
        B, Lq, two_di = xz.shape
        Di, M = two_di // 2, B * Lq
        N = 16
        dev = xz.device
        z = xz[..., Di:]
        f32 = dict(device=dev, dtype=F32)
        g = {}

        # ---- recompute u, delta (checkpoint_lvl = 1, reference :531-534)
        def recompute(cw_, cb_, dtw_, dtb_, dt_, reverse):
            R = dtw_.shape[1]
            u_ = ops.causal_conv1d(xz[..., :Di], mixer._conv_w(cw_), mixer._f32(cb_) if cb_ is not None else None,
                                   silu=True, reverse=reverse)
            delta_ = ops.gemm_tn(dt_, mixer._w(dtw_, act, pad_cols=dt_.shape[1]), k=R,
                                 bias=mixer._f32(dtb_) if dtb_ is not None else None, act=L.ACT_SOFTPLUS,
                                 out_dtype=act if _train16(act, Di) else F32)
            return u_, delta_.view(B, Lq, Di)

        u, delta = recompute(cw, cb, dtw, dtb, dt, False)

        # ---- out_proj backward: d(out_z) = dout @ W_out            (reference :539-540)
        if cfg.has_out:
            dout2 = dout.reshape(M, -1)
            if dout2.dtype != act or dout2.stride(-1) != 1:
                dout2 = dout2.to(act).contiguous()
            dout_z = ops.gemm_tn(dout2, _t(ow, act)).view(B, Lq, Di)
        else:
            dout_z = dout if (dout.dtype == act and dout.stride(-1) == 1) else dout.to(act).contiguous()
            if B > 1 and dout_z.stride(0) != Lq * dout_z.stride(1):
                dout_z = dout_z.contiguous()

        # ---- scan backward, both directions in one launch           (reference :541-561)
        dxz = torch.empty_like(xz)                      # dx | dz written in place (:537-538)
        dz = dxz[..., Di:]
        out_z = torch.empty((B, Lq, Di), device=dev, dtype=act)
        A_f = _A(A, cfg)
        # du / ddelta of the scan (and the x_proj term of d(conv_out)) in the activation dtype when that is 16-bit - what the
        # reference's kernels hand back under autocast (:541-561, :590); the sums over the terms are still formed in fp32
        # (conv backward, sum_cast_colsum).  Needs the specialised backward-scan kernel (AUM_GRAD_16BIT=0: fp32 everywhere).
        g16 = _train16(act, Di)
        gdt = dict(device=dev, dtype=act if g16 else F32)
        # A = -exp(A_log) (module path): the kernel accumulates dA * A = dA_log straight into the A_log gradient when that is a
        # view of the trainer's flat buffer - no zero fill, product and add per direction
        def dA_target(a_param):
            if cfg.a_is_log and _direct(a_param):
                return a_param.grad.view(Di, N), True
            return torch.zeros((Di, N), **f32), False
        dA, dA_direct = dA_target(A)
        dD_buf, dD_direct = _grad_buffer(D, (1, Di)) if D is not None else (None, False)
        du = torch.empty((B, Lq, Di), **gdt)
        ddelta = torch.empty((B, Lq, Di), **gdt)
        dbc = torch.zeros((B, Lq, 2 * N), **f32)
        Dv = mixer._f32(D) if D is not None else None
        d_f = ops.ScanBwdDirection(u, delta, A_f, bc, Dv, du, ddelta, dA, dD_buf.view(Di) if dD_buf is not None else None,
                                   dbc, ck_f, ckpt_valid=True, dA_is_dAlog=dA_direct)
        d_b = None
        dAb_direct = False
        du_b = ddelta_b = None
        if cfg.mode == "v1":
            # each direction writes its own du / ddelta with plain stores (sharing one pair costs a 2 x 100 MB zero fill
            # plus atomic read-modify-writes from both directions at AuM-Base size); their sum is formed where they are
            # consumed anyway: in the conv backward's input and in the cast + column-sum pass of the dt_proj chain
            A_r = _A(A_b, cfg)
            dA_b, dAb_direct = dA_target(A_b)
            du_b = torch.empty((B, Lq, Di), **gdt)
            ddelta_b = torch.empty((B, Lq, Di), **gdt)
            d_b = ops.ScanBwdDirection(u, delta, A_r, bc, Dv, du_b, ddelta_b, dA_b,
                                       dD_buf.view(Di) if dD_buf is not None else None, dbc, ck_b, ckpt_valid=True,
                                       dA_is_dAlog=dAb_direct)
        elif cfg.mode == "v2":
            ub, deltab = recompute(cw_b, cb_b, dtw_b, dtb_b, dt_b, True)
            A_r = _A(A_b, cfg)
            dA_b, dAb_direct = dA_target(A_b)
            dDb_buf, dDb_direct = _grad_buffer(D_b, (1, Di)) if D_b is not None else (None, False)
            du_b = torch.empty((B, Lq, Di), **gdt)
            ddelta_b = torch.empty((B, Lq, Di), **gdt)
            dbc_b = torch.zeros((B, Lq, 2 * N), **f32)
            d_b = ops.ScanBwdDirection(ub, deltab, A_r, bc_b, mixer._f32(D_b) if D_b is not None else None, du_b, ddelta_b,
                                       dA_b, dDb_buf.view(Di) if dDb_buf is not None else None, dbc_b, ck_b, ckpt_valid=True,
                                       dA_is_dAlog=dAb_direct)
        # ddelta comes back already multiplied by softplus'(pre) = 1 - exp(-delta): it IS d(pre-activation)
        ops.selective_scan_bwd(d_f, d_b, z, y_pre, dout_z, dz, out_z, out_scale=cfg.scale, softplus_grad=True)

        # ---- out_proj weight gradient                                          (reference :563-564)
        if cfg.has_out:
            g["ow"] = _wgrad(dout2, out_z.view(M, Di), ow)
            g["ob"] = _deliver_value(ob, dout2.float().sum(0)) if (ob is not None and ob.requires_grad) else None

        # ---- per-branch chain: delta -> dt_proj -> x_proj -> conv                (reference :566-596)
        dxc = torch.empty((B, Lq, Di), device=dev, dtype=act) if cfg.mode == "v2" else None
        dxc_b = torch.empty_like(dxc) if cfg.mode == "v2" else None

        def branch_bwd(sfx, cw_, cb_, xw_, dtw_, dtb_, u_, dt_, du_, ddelta_, dbc_, reverse, dx_out, du2_=None, ddelta2_=None):
            R = dtw_.shape[1]
            # gradient w.r.t. dt_proj's pre-activation (see above): sum of the directions, its 16-bit copy for the two
            # GEMMs below and its column sums (the bias gradient) in ONE pass (:556, :583-586)
            if dtb_ is not None and dtb_.requires_grad:
                db_buf_, db_direct_ = _grad_buffer(dtb_, (1, Di))
            else:
                db_buf_, db_direct_ = None, False
            if Di % 4 == 0:
                dpre_h = ops.sum_cast_colsum(ddelta_.view(M, Di), ddelta2_.view(M, Di) if ddelta2_ is not None else None, act,
                                             db_buf_.view(Di) if db_buf_ is not None else None)
            else:           # odd widths (never AuM's: d_inner = 2 d_model): the same three steps as torch ops
                dpre = ddelta_.view(M, Di) if ddelta2_ is None else ddelta_.view(M, Di) + ddelta2_.view(M, Di)
                if db_buf_ is not None:
                    db_buf_.view(Di).add_(dpre.sum(0))
                dpre_h = dpre.to(act)
            g["dtb" + sfx] = _deliver(dtb_, db_buf_, db_direct_) if db_buf_ is not None else None
            g["dtw" + sfx] = _wgrad(dpre_h, dt_[:, :R], dtw_)                                  # (Di, R)   (:586)
            # dx_dbl = [d(dt) | dB | dC]   (M, R+2N)
            wdb = mixer._round_up(R + 2 * N, 8)
            dxdbl = torch.empty((M, wdb), device=dev, dtype=act)
            ops.gemm_tn(dpre_h, _t(dtw_, act), out=dxdbl[:, :R])                               # d(dt) = dpre @ W_dt
            dxdbl[:, R:R + 2 * N] = dbc_.view(M, 2 * N)
            if wdb > R + 2 * N:
                dxdbl[:, R + 2 * N:] = 0
            g["xw" + sfx] = _wgrad(dxdbl[:, :R + 2 * N], u_.view(M, Di), xw_)                  # (R+2N, Di) (:589)
            # d(conv_out) = du_scan + dx_dbl @ W_x                                              (:590)
            if wdb == xw_.shape[0]:
                wxT = _t(xw_, act)
            else:
                wxT = mixer._cache.get(xw_, f"wT_pad:{act}:{wdb}",
                                       lambda p: torch.nn.functional.pad(p.t().to(act), (0, wdb - p.shape[0])).contiguous())
            du_x = ops.gemm_tn(dxdbl, wxT, out_dtype=du_.dtype)                                # (M, Di), the scan du's dtype
            Wc = cw_.shape[-1]
            dw_buf, dw_direct = _grad_buffer(cw_, (Di, Wc))
            db_buf, db_direct = _grad_buffer(cb_, (1, Di)) if cb_ is not None else (None, False)
            ops.causal_conv1d_bwd(xz[..., :Di], mixer._conv_w(cw_), mixer._f32(cb_) if cb_ is not None else None,
                                  du_.view(B, Lq, Di), dx_out, dw_buf, db_buf.view(Di) if db_buf is not None else None,
                                  silu=True, reverse=reverse, dout2=du_x.view(B, Lq, Di),
                                  dout3=du2_.view(B, Lq, Di) if du2_ is not None else None)   # sums the terms on the fly
            g["cw" + sfx] = _deliver(cw_, dw_buf, dw_direct)
            g["cb" + sfx] = _deliver(cb_, db_buf, db_direct) if cb_ is not None else None

        branch_bwd("", cw, cb, xw, dtw, dtb, u, dt, du, ddelta, dbc, False, dxz[..., :Di] if cfg.mode != "v2" else dxc,
                   du2_=du_b if cfg.mode == "v1" else None, ddelta2_=ddelta_b if cfg.mode == "v1" else None)
        # A = -exp(A_log)  =>  dA_log = dA * A
        g["A"] = None if dA_direct else _deliver_value(A, dA * A_f if cfg.a_is_log else dA)
        g["D"] = _deliver(D, dD_buf, dD_direct) if D is not None else None
        if cfg.mode == "v1":
            g["A_b"] = None if dAb_direct else _deliver_value(A_b, dA_b * A_r if cfg.a_is_log else dA_b)
        elif cfg.mode == "v2":
            branch_bwd("_b", cw_b, cb_b, xw_b, dtw_b, dtb_b, ub, dt_b, du_b, ddelta_b, dbc_b, True, dxc_b)
            g["A_b"] = None if dAb_direct else _deliver_value(A_b, dA_b * A_r if cfg.a_is_log else dA_b)
            g["D_b"] = _deliver(D_b, dDb_buf, dDb_direct) if D_b is not None else None
            torch.add(dxc, dxc_b, out=dxz[..., :Di])

        order = ("cw", "cb", "xw", "dtw", "dtb", "A", "D", "cw_b", "cb_b", "xw_b", "dtw_b", "dtb_b", "A_b", "D_b", "ow", "ob")
        return (None, dxz) + tuple(g.get(k) for k in order)


def _inner_args(m):
    """The 16 tensor arguments of InnerFn from a Mamba module (A_log parameters stand for A: cfg.a_is_log)."""
    v2 = m.bimamba_type == "v2"
    return (m.conv1d.weight, m.conv1d.bias, m.x_proj.weight, m.dt_proj.weight, m.dt_proj.bias, m.A_log, m.D,
            m.conv1d_b.weight if v2 else None, m.conv1d_b.bias if v2 else None, m.x_proj_b.weight if v2 else None,
            m.dt_proj_b.weight if v2 else None, m.dt_proj_b.bias if v2 else None,
            m.A_b_log if m.bimamba_type in ("v1", "v2") else None, m.D_b if v2 else None,
            m.out_proj.weight, m.out_proj.bias)


def mamba_mixer_autograd(module, hidden_states):
    """Mamba.forward under autograd (mamba_simple.py:169-311): in_proj -> fused inner op -> optional layer scale."""
    x = hidden_states
    if torch.is_autocast_enabled():
        x = x.to(torch.get_autocast_dtype("cuda"))
    B, Lq, Dm = x.shape
    x2 = x.reshape(B * Lq, Dm)
    if x2.stride(-1) != 1:
        x2 = x2.contiguous()
    xz = LinearFn.apply(x2, module.in_proj.weight, module.in_proj.bias, None).view(B, Lq, 2 * module.d_inner)
    scale = 0.5 if (module.bimamba_type == "v2" and module.if_devide_out) else 1.0
    cfg = InnerCfg(module.bimamba_type, True, scale, True)
    out = InnerFn.apply(cfg, xz, *_inner_args(module))
    gamma = getattr(module, "gamma", None)
    if getattr(module, "init_layer_scale", None) is not None and gamma is not None:
        out = out * gamma.to(out.dtype)
    return out


# ------------------------------------------------------------------------------------------------------
class ScanFn(torch.autograd.Function):
    """selective_scan_fn under autograd (SelectiveScanFn, selective_scan_interface.py:14-74), one direction.
    Token-major: u, z (B, L, D) activation dtype; delta (B, L, D) fp32, FINAL (bias and softplus already applied by
    differentiable torch ops in the caller); A (D, 16) fp32; bc (B, L, 32) fp32 packed [B|C]; D (D,) fp32 or None."""

    @staticmethod
    def forward(ctx, u, delta, A, bc, Dv, z):
        B, Lq, Dch = u.shape
        N = A.shape[1]
        ck = ops.scan_bwd_workspace(B, Lq, Dch, u.device)
        out = torch.empty_like(u)
        y_pre = torch.empty_like(u) if z is not None else None
        d = ops.ScanDirection(u, delta, A, bc[..., :N], bc[..., N:], Dv, ckpt=ck)
        ops.selective_scan(d, None, z, out=out, y_pre=y_pre)
        ctx.save_for_backward(u, delta, A, bc, Dv, z, y_pre if y_pre is not None else out, ck)
        ctx.has_z, ctx.has_D = z is not None, Dv is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        u, delta, A, bc, Dv, z, y_pre, ck = ctx.saved_tensors
        B, Lq, Dch = u.shape
        N = A.shape[1]
        f32 = dict(device=u.device, dtype=F32)
        dout = dout if (dout.dtype == u.dtype and dout.is_contiguous()) else dout.to(u.dtype).contiguous()
        if not ctx.has_z:
            z = None                                  # no gate: out = y, the kernel skips dz / out_z
        du = torch.empty((B, Lq, Dch), **f32)
        dd = torch.empty((B, Lq, Dch), **f32)
        dbc = torch.zeros((B, Lq, 2 * N), **f32)
        dA = torch.zeros((Dch, N), **f32)
        dD = torch.zeros((Dch,), **f32) if ctx.has_D else None
        dz = torch.empty_like(u) if ctx.has_z else None
        oz = torch.empty_like(u) if ctx.has_z else None
        d = ops.ScanBwdDirection(u, delta, A, bc, Dv if ctx.has_D else None, du, dd, dA, dD, dbc, ck, ckpt_valid=True)
        ops.selective_scan_bwd(d, None, z, y_pre, dout, dz, oz)
        return du.to(u.dtype), dd, dA, dbc, dD, dz


class ConvFn(torch.autograd.Function):
    """causal_conv1d_fn under autograd (pip causal_conv1d's CausalConv1dFn; call sites selective_scan_interface.py
    :646,:683 and mamba_simple.py:275).  Token-major x (B, L, D)."""

    @staticmethod
    def forward(ctx, x, weight, bias, silu):
        ctx.save_for_backward(x)
        ctx.weight, ctx.bias, ctx.silu = weight, bias, silu
        return ops.causal_conv1d(x, mixer._conv_w(weight), mixer._f32(bias) if bias is not None else None, silu=silu)

    @staticmethod
    def backward(ctx, dout):
        (x,) = ctx.saved_tensors
        w, b = ctx.weight, ctx.bias
        B, Lq, D = x.shape
        dx = torch.empty((B, Lq, D), device=x.device, dtype=x.dtype)
        dw_buf, dw_direct = _grad_buffer(w, (D, w.shape[-1]))
        db_buf, db_direct = _grad_buffer(b, (1, D)) if b is not None else (None, False)
        ops.causal_conv1d_bwd(x, mixer._conv_w(w), mixer._f32(b) if b is not None else None, dout.float().contiguous(),
                              dx, dw_buf, db_buf.view(D) if db_buf is not None else None, silu=ctx.silu)
        return dx, _deliver(w, dw_buf, dw_direct), _deliver(b, db_buf, db_direct) if b is not None else None, None


# ------------------------------------------------------------------------------------------------------
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
        ctx.save_for_backward(res, rstd)
        ctx.weight = weight
        ctx.has_res, ctx.has_bias, ctx.prenorm = residual is not None, bias is not None, prenorm
        ctx.x_dtype, ctx.r_dtype = x.dtype, (residual.dtype if residual is not None else None)
        y = y.view(shape)
        return (y, res.view(shape)) if prenorm else y

    @staticmethod
    def backward(ctx, dy, *args):
        res, rstd = ctx.saved_tensors
        weight = ctx.weight
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
            dw_buf, dw_direct = _grad_buffer(weight, (1, dim))
            dx, dri = ops.add_rmsnorm_bwd(dy2, dro2, res, rstd, mixer._f32(weight), dw_buf.view(dim), want_dres_in=ctx.has_res)
            return (dx.view(dy.shape), _deliver(weight, dw_buf, dw_direct), None,
                    dri.view(dy.shape) if dri is not None else None, None, None, None)
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
