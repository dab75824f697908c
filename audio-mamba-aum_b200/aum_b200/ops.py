"""Token-major op wrappers over the C ABI (include/aum_b200.h).

Every tensor here is "token-major": shape (rows, C) or (B, L, C) with the channel axis contiguous and a row
pitch (leading dimension) that may exceed C (views into wider buffers are fine and never copied).
PyTorch allocates; the library only launches kernels on the current stream.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib as L

# bench.py hook: when set to a list, (name, start_event, end_event) is appended around each selective-scan launch
PROFILE = None


def _as_rows(t: torch.Tensor) -> Tuple[int, int, int]:
    """(rows, cols, ld) of a token-major tensor (2-D, or 3-D whose first two axes collapse)."""
    if t.stride(-1) != 1 and t.shape[-1] != 1:
        raise L.AumError("token-major tensor must have a contiguous last axis")
    if t.dim() == 2:
        return t.shape[0], t.shape[1], (t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1]))
    if t.dim() == 3:
        B, Lq, Cc = t.shape
        ld = t.stride(1) if Lq > 1 else max(t.stride(1), Cc)
        if B > 1 and t.stride(0) != Lq * ld:
            raise L.AumError("3-D token-major tensor must collapse to rows (stride(0) == L*stride(1))")
        return B * Lq, Cc, ld
    raise L.AumError("expected a 2-D or 3-D tensor")


def gemm_tn(a: torch.Tensor, w: torch.Tensor, *, out: Optional[torch.Tensor] = None,
            out_dtype: Optional[torch.dtype] = None, out2: Optional[torch.Tensor] = None, split: Optional[int] = None,
            bias: Optional[torch.Tensor] = None, row_scale: Optional[torch.Tensor] = None,
            act: int = L.ACT_NONE, backend: int = L.GEMM_AUTO, k: Optional[int] = None) -> torch.Tensor:
    """out[M,N] = act(row_scale * (a[M,K] @ w[N,K]^T) + bias).  a, w: same dtype, K-contiguous rows.
    ``k`` restricts the reduction to the first k columns of both operands (zero-padded weight copies)."""
    L.require_cuda(a, w)
    M, Ka, lda = _as_rows(a)
    N, Kw, ldw = _as_rows(w)
    K = k if k is not None else Ka
    if K > Ka or K > Kw or a.dtype != w.dtype:
        raise L.AumError(f"gemm_tn: incompatible operands a{tuple(a.shape)} {a.dtype} w{tuple(w.shape)} {w.dtype} K={K}")
    if out2 is None:
        split_ = N
        if out is None:
            out = torch.empty((M, N), device=a.device, dtype=out_dtype or a.dtype)
        ncols = N
    else:
        split_ = int(split)
        ncols = split_
        if out is None:
            raise L.AumError("gemm_tn: split output needs explicit out and out2")
    Mo, No, ldc = _as_rows(out)
    if Mo != M or No < ncols:
        raise L.AumError("gemm_tn: bad output shape")
    ldc2, c2dt, p2 = 0, L.F32, None
    if out2 is not None:
        M2, N2, ldc2 = _as_rows(out2)
        if M2 != M or N2 < N - split_:
            raise L.AumError("gemm_tn: bad second output shape")
        c2dt, p2 = L.dt(out2.dtype), L.ptr(out2)
    if bias is not None and (bias.dtype != torch.float32 or bias.numel() != N):
        raise L.AumError("gemm_tn: bias must be fp32 of length N")
    if row_scale is not None and (row_scale.dtype != torch.float32 or row_scale.numel() != M):
        raise L.AumError("gemm_tn: row_scale must be fp32 of length M")
    rc = L.lib().aum_gemm_tn(L.ptr(a), lda, L.ptr(w), ldw, L.dt(a.dtype),
                             L.ptr(out), ldc, L.dt(out.dtype), p2, ldc2, c2dt, split_,
                             M, N, K, L.ptr(bias), L.ptr(row_scale), act, backend, L.stream())
    L.check(rc, "aum_gemm_tn")
    return out


def gemm_wgrad(dy: torch.Tensor, x: torch.Tensor, dw: torch.Tensor) -> torch.Tensor:
    """dw[No, Ki] += dy[T, No]^T @ x[T, Ki]: the weight-gradient reduction over the token axis (aum_gemm_wgrad).
    dy, x: token-major 2-D (views into wider buffers are fine); dw: fp32, accumulated into (row pitch >= Ki).
    16-bit operands run the tcgen05 MN-major kernel; fp32 operands (strict-parity tier) are transposed once each and
    go through the fp32 CUDA-core GEMM, whose result is added to dw."""
    L.require_cuda(dy, x, dw)
    T, No, ld_dy = _as_rows(dy)
    T2, Ki, ld_x = _as_rows(x)
    if T2 != T or dy.dtype != x.dtype or dw.dtype != torch.float32 or dw.dim() != 2 or tuple(dw.shape) != (No, Ki):
        raise L.AumError(f"gemm_wgrad: incompatible operands dy{tuple(dy.shape)} x{tuple(x.shape)} dw{tuple(dw.shape)} {dw.dtype}")
    if dw.stride(1) != 1 and Ki != 1:
        raise L.AumError("gemm_wgrad: dw rows must be contiguous")
    if T == 0:
        return dw
    ok16 = lambda t, ld: t.data_ptr() % 16 == 0 and (ld * 2) % 16 == 0
    if dy.dtype in (torch.float16, torch.bfloat16) and ok16(dy, ld_dy) and ok16(x, ld_x):
        rc = L.lib().aum_gemm_wgrad(L.ptr(dy), ld_dy, L.ptr(x), ld_x, L.dt(dy.dtype), L.ptr(dw),
                                    dw.stride(0) if No > 1 else max(dw.stride(0), Ki), T, No, Ki, L.stream())
        L.check(rc, "aum_gemm_wgrad")
        return dw
    # fp32 tier / unaligned views: dW = (dy^T)[No, T] @ (x^T)[Ki, T]^T through the K-contiguous GEMM
    dyT = transpose(dy.reshape(1, T, No) if dy.is_contiguous() else dy.contiguous().view(1, T, No)).view(No, T)
    xT = transpose(x.reshape(1, T, Ki) if x.is_contiguous() else x.contiguous().view(1, T, Ki)).view(Ki, T)
    part = gemm_tn(dyT, xT, out_dtype=torch.float32)
    dw.add_(part)
    return dw


def causal_conv1d(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], *, silu: bool = True,
                  reverse: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x: (B, L, D) token-major (may be a channel slice of a wider buffer); w: (D, W) fp32; bias (D) fp32."""
    L.require_cuda(x, w)
    B, Lq, D = x.shape
    _, _, ldx = _as_rows(x)
    if w.dtype != torch.float32 or not w.is_contiguous() or w.shape[0] != D:
        raise L.AumError("causal_conv1d: weight must be contiguous fp32 (D, W)")
    if bias is not None and (bias.dtype != torch.float32 or not bias.is_contiguous()):
        raise L.AumError("causal_conv1d: bias must be contiguous fp32")
    if out is None:
        out = torch.empty((B, Lq, D), device=x.device, dtype=x.dtype)
    _, _, ldo = _as_rows(out)
    rc = L.lib().aum_causal_conv1d_fwd(L.ptr(x), ldx, L.ptr(w), L.ptr(bias), L.ptr(out), ldo,
                                       B, Lq, D, w.shape[1], L.dt(x.dtype), int(silu), int(reverse), L.stream())
    L.check(rc, "aum_causal_conv1d_fwd")
    return out


def conv_xproj_eligible(x: torch.Tensor, conv_w: torch.Tensor, wx: torch.Tensor, R: int, N2: int) -> bool:
    """Shapes / dtypes / alignments the fused conv + x_proj kernel takes (everything AuM-Small and AuM-Base produce)."""
    if x.dtype not in (torch.float16, torch.bfloat16) or wx.dtype != x.dtype or conv_w.shape[-1] != 4:
        return False
    Di = x.shape[-1]
    ldx = _as_rows(x)[2]
    return (R % 8 == 0 and R + N2 <= 128 and Di % 8 == 0 and x.data_ptr() % 16 == 0 and (ldx * 2) % 16 == 0
            and wx.is_contiguous() and wx.data_ptr() % 16 == 0)


def conv_xproj(x: torch.Tensor, conv_w: torch.Tensor, conv_b: Optional[torch.Tensor], wx: torch.Tensor, R: int,
               dt: torch.Tensor, bc: torch.Tensor, *, reverse: bool = False, u: Optional[torch.Tensor] = None) -> torch.Tensor:
    """u = silu(causal_conv1d(x) + b) and u @ wx^T -> dt[:, :R] (16-bit) | bc (fp32) in ONE launch (aum_conv_xproj_fwd).
    x: (B, L, Di) token-major view; conv_w (Di, 4) fp32; wx (R + N2, Di) in x's dtype; returns u (B, L, Di)."""
    L.require_cuda(x, conv_w, wx, dt, bc)
    B, Lq, Di = x.shape
    ldx = _as_rows(x)[2]
    N2 = wx.shape[0] - R
    if u is None:
        u = torch.empty((B, Lq, Di), device=x.device, dtype=x.dtype)
    if conv_w.dtype != torch.float32 or not conv_w.is_contiguous() or tuple(conv_w.shape) != (Di, 4):
        raise L.AumError("conv_xproj: conv weight must be contiguous fp32 (Di, 4)")
    if bc.dtype != torch.float32 or dt.dtype != x.dtype or bc.shape[-1] < N2 or dt.shape[-1] < R:
        raise L.AumError("conv_xproj: dt must have x's dtype and >= R columns, bc fp32 with >= N2 columns")
    rc = L.lib().aum_conv_xproj_fwd(L.ptr(x), ldx, L.ptr(conv_w), L.ptr(conv_b), L.ptr(wx), wx.stride(0),
                                    L.ptr(u), _as_rows(u)[2], L.ptr(dt), _as_rows(dt)[2], L.ptr(bc), _as_rows(bc)[2],
                                    B, Lq, Di, R, N2, L.dt(x.dtype), int(reverse), L.stream())
    L.check(rc, "aum_conv_xproj_fwd")
    return u


class ScanDirection:
    """One time direction of the scan (struct aum_scan_dir).  All tensors token-major:
    u, delta: (B, L, D); A: (D, N) fp32; Bm, Cm: (B, L, N); D, delta_bias: (D,) fp32 or None."""

    def __init__(self, u, delta, A, Bm, Cm, D=None, delta_bias=None, delta_softplus=False, last_state=None, ckpt=None):
        self.u, self.delta, self.A, self.Bm, self.Cm = u, delta, A, Bm, Cm
        self.D, self.delta_bias, self.delta_softplus, self.last_state = D, delta_bias, delta_softplus, last_state
        self.ckpt = ckpt       # optional fp32 workspace (scan_bwd_workspace) receiving the state checkpoints

    def _struct(self, B, Lq, Dch, N):
        L.require_cuda(self.u, self.delta, self.A, self.Bm, self.Cm)
        for t, name in ((self.A, "A"), (self.D, "D"), (self.delta_bias, "delta_bias")):
            if t is not None and (t.dtype != torch.float32 or not t.is_contiguous()):
                raise L.AumError(f"scan: {name} must be contiguous fp32")
        if tuple(self.u.shape) != (B, Lq, Dch) or tuple(self.delta.shape) != (B, Lq, Dch):
            raise L.AumError("scan: u/delta shape mismatch")
        if tuple(self.Bm.shape) != (B, Lq, N) or tuple(self.Cm.shape) != (B, Lq, N) or self.Bm.dtype != self.Cm.dtype:
            raise L.AumError("scan: B/C shape or dtype mismatch")
        if tuple(self.A.shape) != (Dch, N):
            raise L.AumError("scan: A must be (D, N)")
        s = L.ScanDir()
        s.u, s.ld_u = self.u.data_ptr(), _as_rows(self.u)[2]
        s.delta, s.ld_delta, s.delta_dtype = self.delta.data_ptr(), _as_rows(self.delta)[2], L.dt(self.delta.dtype)
        s.A = self.A.data_ptr()
        s.Bm, s.ld_B = self.Bm.data_ptr(), _as_rows(self.Bm)[2]
        s.Cm, s.ld_C = self.Cm.data_ptr(), _as_rows(self.Cm)[2]
        s.bc_dtype = L.dt(self.Bm.dtype)
        s.D = self.D.data_ptr() if self.D is not None else None
        s.delta_bias = self.delta_bias.data_ptr() if self.delta_bias is not None else None
        s.delta_softplus = int(bool(self.delta_softplus))
        s.last_state = self.last_state.data_ptr() if self.last_state is not None else None
        s.ckpt = self.ckpt.data_ptr() if self.ckpt is not None else None
        return s


def selective_scan(fwd: Optional[ScanDirection], bwd: Optional[ScanDirection], z: Optional[torch.Tensor], *,
                   out: Optional[torch.Tensor] = None, out_scale: float = 1.0,
                   y_pre: Optional[torch.Tensor] = None, z_pregated: bool = False) -> torch.Tensor:
    """out = out_scale * (y_fwd + y_bwd) * silu(z); either direction may be None.  Token-major (B, L, D).
    y_pre (optional, same shape/dtype/pitch as out) receives the pre-gate sum y_fwd + y_bwd for the backward pass.
    z_pregated: z already holds silu(z) (the in_proj GEMM epilogue applied it)."""
    ref = fwd if fwd is not None else bwd
    if ref is None:
        raise L.AumError("selective_scan: no direction given")
    B, Lq, Dch = ref.u.shape
    N = ref.A.shape[1]
    if out is None:
        out = torch.empty((B, Lq, Dch), device=ref.u.device, dtype=ref.u.dtype)
    if out.dtype != ref.u.dtype or (z is not None and z.dtype != ref.u.dtype):
        raise L.AumError("selective_scan: u, z and out must share a dtype")
    import ctypes as C
    sf = fwd._struct(B, Lq, Dch, N) if fwd is not None else None
    sb = bwd._struct(B, Lq, Dch, N) if bwd is not None else None
    L.require_cuda(ref.u, out, z, y_pre, fwd.u if fwd is not None else None, bwd.u if bwd is not None else None)
    ldz = _as_rows(z)[2] if z is not None else 0
    if PROFILE is not None:
        ev0 = torch.cuda.Event(enable_timing=True)
        ev0.record()
    rc = L.lib().aum_selective_scan_fwd(C.byref(sf) if sf is not None else None,
                                        C.byref(sb) if sb is not None else None,
                                        L.ptr(z), ldz, L.ptr(out), _as_rows(out)[2],
                                        B, Lq, Dch, N, L.dt(out.dtype), float(out_scale),
                                        L.ptr(y_pre), _as_rows(y_pre)[2] if y_pre is not None else 0,
                                        L.SCAN_Z_PREGATED if z_pregated else 0, L.stream())
    L.check(rc, "aum_selective_scan_fwd")
    if PROFILE is not None:
        ev1 = torch.cuda.Event(enable_timing=True)
        ev1.record()
        PROFILE.append(("selective_scan", ev0, ev1))
    return out


def add_rmsnorm(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None,
                residual: Optional[torch.Tensor] = None, *, eps: float = 1e-5, prenorm: bool = False,
                residual_dtype: Optional[torch.dtype] = torch.float32, out_dtype: Optional[torch.dtype] = None,
                rstd: Optional[torch.Tensor] = None):
    """y = rmsnorm(x + residual) * weight (+ bias); with prenorm also returns the new residual (x + residual)."""
    L.require_cuda(x, weight)
    rows, dim, ldx = _as_rows(x)
    if weight.dtype != torch.float32 or not weight.is_contiguous():
        raise L.AumError("add_rmsnorm: weight must be contiguous fp32")
    y = torch.empty(x.shape, device=x.device, dtype=out_dtype or x.dtype)
    res_out = None
    if prenorm:
        rd = residual.dtype if residual is not None else (residual_dtype or x.dtype)
        res_out = torch.empty(x.shape, device=x.device, dtype=rd)
    ldr = _as_rows(residual)[2] if residual is not None else 0
    rc = L.lib().aum_add_rmsnorm_fwd(L.ptr(x), ldx, L.dt(x.dtype),
                                     L.ptr(residual), ldr, L.dt(residual.dtype) if residual is not None else L.F32,
                                     L.ptr(weight), L.ptr(bias),
                                     L.ptr(y), _as_rows(y)[2], L.dt(y.dtype),
                                     L.ptr(res_out), _as_rows(res_out)[2] if res_out is not None else 0,
                                     L.dt(res_out.dtype) if res_out is not None else L.F32,
                                     L.ptr(rstd), rows, dim, float(eps), L.stream())
    L.check(rc, "aum_add_rmsnorm_fwd")
    return (y, res_out) if prenorm else y


def transpose(src: torch.Tensor, dst: Optional[torch.Tensor] = None, dst_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """(B, R, C) -> (B, C, R) with optional dtype change; src rows must be contiguous along C."""
    L.require_cuda(src)
    B, R, Cc = src.shape
    if src.stride(2) != 1 and Cc != 1:
        raise L.AumError("transpose: source last axis must be contiguous")
    if dst is None:
        dst = torch.empty((B, Cc, R), device=src.device, dtype=dst_dtype or src.dtype)
    if dst.stride(2) != 1 and R != 1:
        raise L.AumError("transpose: destination last axis must be contiguous")
    rc = L.lib().aum_transpose(L.ptr(src), src.stride(0), src.stride(1), L.ptr(dst), dst.stride(0), dst.stride(1),
                               B, R, Cc, L.dt(src.dtype), L.dt(dst.dtype), L.stream())
    L.check(rc, "aum_transpose")
    return dst


def causal_conv1d_bwd(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], dout: torch.Tensor,
                      dx: torch.Tensor, dw: torch.Tensor, dbias: Optional[torch.Tensor], *, silu: bool = True,
                      reverse: bool = False, dout2: Optional[torch.Tensor] = None,
                      dout3: Optional[torch.Tensor] = None) -> None:
    """Backward of causal_conv1d.  x, dx: (B, L, D) token-major (dtype); dout (+ dout2 + dout3, summed on the fly):
    (B, L, D), all fp32 or all of x's 16-bit dtype; dw (D, W) / dbias (D) fp32 are accumulated into."""
    L.require_cuda(x, w, dout, dx, dw)
    B, Lq, D = x.shape
    if dout.dtype not in (torch.float32, x.dtype) or dw.dtype != torch.float32 or dx.dtype != x.dtype:
        raise L.AumError("causal_conv1d_bwd: dout must be fp32 or match x, dw must be fp32 and dx must match x")
    for extra in (dout2, dout3):
        if extra is not None and (extra.dtype != dout.dtype or _as_rows(extra)[2] != _as_rows(dout)[2]):
            raise L.AumError("causal_conv1d_bwd: dout2 / dout3 must have dout's dtype and pitch")
    rc = L.lib().aum_causal_conv1d_bwd(L.ptr(x), _as_rows(x)[2], L.ptr(w), L.ptr(bias), L.ptr(dout), L.ptr(dout2), L.ptr(dout3),
                                       _as_rows(dout)[2], L.dt(dout.dtype),
                                       L.ptr(dx), _as_rows(dx)[2], L.ptr(dw), L.ptr(dbias), B, Lq, D, w.shape[1],
                                       L.dt(x.dtype), int(silu), int(reverse), L.stream())
    L.check(rc, "aum_causal_conv1d_bwd")


def sum_cast_colsum(a: torch.Tensor, b: Optional[torch.Tensor], out_dtype: torch.dtype,
                    colsum: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = (a + b).to(out_dtype) and colsum += (a + b).sum(0) in one pass (aum_sum_cast_colsum).  a, b: (rows, cols) of
    one dtype (fp32 or 16-bit); the sum is formed in fp32."""
    L.require_cuda(a, b, colsum)
    rows, cols, ld = _as_rows(a)
    if b is not None and (b.dtype != a.dtype or _as_rows(b) != (rows, cols, ld)):
        raise L.AumError("sum_cast_colsum: a and b must share dtype, shape and pitch")
    if colsum is not None and (colsum.dtype != torch.float32 or colsum.numel() != cols or not colsum.is_contiguous()):
        raise L.AumError("sum_cast_colsum: colsum must be a contiguous fp32 vector of `cols` elements")
    al = 4 * a.element_size()
    if cols % 4 != 0 or ld % 4 != 0 or a.data_ptr() % al != 0 or (b is not None and b.data_ptr() % al != 0):
        raise L.AumError("sum_cast_colsum: rows must be addressable as 4-element vectors (cols % 4 == 0, 16-byte aligned)")
    out = torch.empty((rows, cols), device=a.device, dtype=out_dtype)
    rc = L.lib().aum_sum_cast_colsum(L.ptr(a), L.ptr(b), ld, L.dt(a.dtype), L.ptr(out), cols, L.dt(out_dtype), L.ptr(colsum), rows, cols,
                                     L.stream())
    L.check(rc, "aum_sum_cast_colsum")
    return out


class ScanBwdDirection:
    """One time direction of the scan backward (struct aum_scan_bwd_dir).  u: (B,L,D) dtype; delta: (B,L,D) fp32
    post-softplus; A: (D,16) fp32; bc: (B,L,32) fp32 packed [B|C]; outputs du, ddelta (B,L,D) fp32 - or both of u's
    16-bit dtype (training configuration only, see aum_b200.h) -, dA (D,16), dD (D), dbc (B,L,32) fp32 accumulated into;
    ckpt: fp32 workspace."""

    def __init__(self, u, delta, A, bc, D, du, ddelta, dA, dD, dbc, ckpt, ckpt_valid=False, dA_is_dAlog=False):
        self.t = (u, delta, A, bc, D, du, ddelta, dA, dD, dbc, ckpt)
        self.ckpt_valid = ckpt_valid   # ckpt was filled by selective_scan(... ScanDirection(ckpt=...)) of the same call shape
        self.dA_is_dAlog = dA_is_dAlog   # dA receives dA * A: the gradient w.r.t. A_log of A = -exp(A_log)

    def _struct(self):
        u, delta, A, bc, D, du, ddelta, dA, dD, dbc, ckpt = self.t
        for t_ in (A, bc, dA, dbc, ckpt):
            if t_.dtype != torch.float32:
                raise L.AumError("scan bwd: A/bc/dA/dbc/ckpt must be fp32")
        if delta.dtype not in (torch.float32, u.dtype):
            raise L.AumError("scan bwd: delta must be fp32 or of u's dtype")
        if du.dtype != ddelta.dtype or du.dtype not in (torch.float32, u.dtype):
            raise L.AumError("scan bwd: du and ddelta must share a dtype: fp32 or u's")
        s = L.ScanBwdDir()
        s.u, s.ld_u = u.data_ptr(), _as_rows(u)[2]
        s.delta, s.ld_delta = delta.data_ptr(), _as_rows(delta)[2]
        s.A = A.data_ptr()
        s.BC, s.ld_bc = bc.data_ptr(), _as_rows(bc)[2]
        s.D = D.data_ptr() if D is not None else None
        s.du, s.ld_du = du.data_ptr(), _as_rows(du)[2]
        s.ddelta, s.ld_dd = ddelta.data_ptr(), _as_rows(ddelta)[2]
        s.dA = dA.data_ptr()
        s.dD = dD.data_ptr() if dD is not None else None
        s.dBC, s.ld_dbc = dbc.data_ptr(), _as_rows(dbc)[2]
        s.ckpt = ckpt.data_ptr()
        s.ckpt_valid = int(bool(self.ckpt_valid))
        s.dgrad_dtype = L.dt(du.dtype)
        s.delta_dtype = L.dt(delta.dtype)
        s.dA_is_dAlog = int(bool(self.dA_is_dAlog))
        B, Lq, Dch = u.shape
        n = L.lib().aum_selective_scan_bwd_dbc_ws_floats(B, Lq, Dch)
        self._ws = torch.empty(n, device=u.device, dtype=torch.float32)   # per-warp dB|dC partials (kept alive here)
        s.dbc_ws = self._ws.data_ptr()
        return s


def scan_bwd_workspace(batch: int, Lq: int, D: int, device) -> torch.Tensor:
    n = L.lib().aum_selective_scan_bwd_workspace_floats(batch, Lq, D)
    return torch.empty(n, device=device, dtype=torch.float32)


def selective_scan_bwd(fwd: Optional[ScanBwdDirection], bwd: Optional[ScanBwdDirection], z, y_pre, dout, dz, out_z,
                       *, out_scale: float = 1.0, softplus_grad: bool = False) -> None:
    """softplus_grad: the ddelta outputs are multiplied by (1 - exp(-delta)), i.e. they become the gradient w.r.t. the
    pre-softplus dt_proj output."""
    import ctypes as C
    ref = fwd if fwd is not None else bwd
    u = ref.t[0]
    B, Lq, Dch = u.shape
    L.require_cuda(u, z, y_pre, dout, dz, out_z, *(t_ for d_ in (fwd, bwd) if d_ is not None for t_ in d_.t))
    sf = fwd._struct() if fwd is not None else None
    sb = bwd._struct() if bwd is not None else None
    ld = lambda t: _as_rows(t)[2] if t is not None else 0
    rc = L.lib().aum_selective_scan_bwd(C.byref(sf) if sf is not None else None, C.byref(sb) if sb is not None else None,
                                        L.ptr(z), ld(z), L.ptr(y_pre), ld(y_pre), L.ptr(dout), ld(dout),
                                        L.ptr(dz), ld(dz), L.ptr(out_z), ld(out_z),
                                        B, Lq, Dch, 16, L.dt(u.dtype), float(out_scale), int(softplus_grad), L.stream())
    L.check(rc, "aum_selective_scan_bwd")


def add_rmsnorm_bwd(dy: torch.Tensor, dres_out: Optional[torch.Tensor], r: torch.Tensor, rstd: torch.Tensor,
                    weight: torch.Tensor, dweight: torch.Tensor, *, want_dres_in: bool):
    """Backward of add_rmsnorm (RMSNorm, no bias).  dy: (rows, dim) dtype; dres_out: (rows, dim) fp32 or None;
    r: saved residual_out fp32; returns (dx [dy.dtype], dres_in [fp32] or None); dweight (dim) fp32 accumulated into."""
    L.require_cuda(dy, r, rstd, weight, dweight)
    rows, dim, ld_dy = _as_rows(dy)
    dx = torch.empty((rows, dim), device=dy.device, dtype=dy.dtype)
    dri = torch.empty((rows, dim), device=dy.device, dtype=torch.float32) if want_dres_in else None
    rc = L.lib().aum_add_rmsnorm_bwd(L.ptr(dy), ld_dy, L.dt(dy.dtype), L.ptr(dres_out),
                                     _as_rows(dres_out)[2] if dres_out is not None else 0,
                                     L.ptr(r), _as_rows(r)[2], L.ptr(rstd), L.ptr(weight),
                                     L.ptr(dx), dim, L.ptr(dri), dim, L.ptr(dweight), rows, dim, L.stream())
    L.check(rc, "aum_add_rmsnorm_bwd")
    return dx, dri


def adam_step(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, *, lr: float, betas=(0.9, 0.999),
              eps: float = 1e-8, weight_decay: float = 0.0, step: int = 1, grad_scale: float = 1.0) -> None:
    """In-place torch.optim.Adam update of the flat fp32 buffer p from its flat gradient g and moments m, v
    (aum_adam_step; reference: src/traintest.py:32-34,169)."""
    L.require_cuda(p, g, m, v)
    for t in (p, g, m, v):
        if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != p.numel():
            raise L.AumError("adam_step: p, g, m, v must be contiguous fp32 buffers of one size")
    rc = L.lib().aum_adam_step(L.ptr(p), L.ptr(g), L.ptr(m), L.ptr(v), p.numel(), float(lr), float(betas[0]),
                               float(betas[1]), float(eps), float(weight_decay), int(step), float(grad_scale), L.stream())
    L.check(rc, "aum_adam_step")


def adam_step_dev(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, step_dev: torch.Tensor, *, lr: float,
                  betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0, grad_scale: float = 1.0,
                  p16: Optional[torch.Tensor] = None) -> None:
    """adam_step with the step number in device memory (step_dev: int32 scalar tensor, incremented by the call) and an
    optional 16-bit shadow copy of the updated parameters (aum_adam_step_dev): every argument of the launch is then
    step-independent, so a captured CUDA graph of the training step can be replayed."""
    L.require_cuda(p, g, m, v, step_dev, p16)
    for t in (p, g, m, v):
        if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != p.numel():
            raise L.AumError("adam_step_dev: p, g, m, v must be contiguous fp32 buffers of one size")
    if step_dev.dtype != torch.int32 or step_dev.numel() != 1:
        raise L.AumError("adam_step_dev: step_dev must be an int32 scalar tensor")
    if p16 is not None and (p16.dtype not in (torch.float16, torch.bfloat16) or p16.numel() != p.numel() or not p16.is_contiguous()):
        raise L.AumError("adam_step_dev: p16 must be a contiguous fp16 / bf16 buffer of p's size")
    rc = L.lib().aum_adam_step_dev(L.ptr(p), L.ptr(g), L.ptr(m), L.ptr(v), p.numel(), float(lr), float(betas[0]),
                                   float(betas[1]), float(eps), float(weight_decay), L.ptr(step_dev), float(grad_scale),
                                   L.ptr(p16), L.dt(p16.dtype) if p16 is not None else L.BF16, L.stream())
    L.check(rc, "aum_adam_step_dev")
    L._launches += 1      # two kernels: the counter increment and the update


def patchify(x: torch.Tensor, patch, out_dtype: torch.dtype) -> torch.Tensor:
    """(B, T, F) fp32 spectrogram -> (B * gf * gt, pf * pt) im2col rows of the stride == kernel patch conv
    (aum_patchify; reference: tokenization.py:278-310 via mamba_models.py:510-515)."""
    L.require_cuda(x)
    if x.dtype != torch.float32 or not x.is_contiguous() or x.dim() != 3:
        raise L.AumError("patchify: x must be a contiguous fp32 (B, T, F) tensor")
    B, T_, F_ = x.shape
    pf, pt = patch
    if F_ % pf or T_ % pt or pt % 4:
        raise L.AumError("patchify: patch must tile the spectrogram and pt must be a multiple of 4")
    cols = torch.empty((B * (F_ // pf) * (T_ // pt), pf * pt), device=x.device, dtype=out_dtype)
    rc = L.lib().aum_patchify(L.ptr(x), L.ptr(cols), B, T_, F_, pf, pt, L.dt(out_dtype), L.stream())
    L.check(rc, "aum_patchify")
    return cols


def assemble_tokens(tok: torch.Tensor, pos: torch.Tensor, cls: torch.Tensor) -> torch.Tensor:
    """tok (B, N, Dm) fp32, pos (N + 1, Dm) fp32 (slot 0 = cls), cls (Dm) fp32 -> hidden (B, N + 1, Dm) fp32 with the
    cls token in the middle (aum_assemble_tokens; reference: mamba_models.py:525-541)."""
    L.require_cuda(tok, pos, cls)
    B, N, Dm = tok.shape
    for t in (tok, pos, cls):
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise L.AumError("assemble_tokens: contiguous fp32 tensors required")
    if pos.numel() != (N + 1) * Dm or cls.numel() != Dm:
        raise L.AumError("assemble_tokens: pos must be (N + 1, Dm) and cls (Dm)")
    hidden = torch.empty((B, N + 1, Dm), device=tok.device, dtype=torch.float32)
    rc = L.lib().aum_assemble_tokens(L.ptr(tok), L.ptr(pos), L.ptr(cls), L.ptr(hidden), B, N, Dm, L.stream())
    L.check(rc, "aum_assemble_tokens")
    return hidden
