"""Host-side orchestration of one Mamba mixer block on the B200 engine.

Reference path being replaced: Mamba.forward fast path
(/root/reference/vim-mamba_ssm/mamba_ssm/modules/mamba_simple.py:169-311) ->
bimamba_inner_fn / mamba_inner_fn / mamba_inner_fn_no_out_proj
(vim-mamba_ssm/mamba_ssm/ops/selective_scan_interface.py:437-517, 292-365, 155-224).

Data layout in HBM (all token-major, rows = B*L tokens):
    hidden (M, Dm) --in_proj--> xz (M, 2Di)  [x | z]
    x --conv+SiLU--> u (M, Di)
    u --x_proj--> dt (M, Rpad) act dtype  +  bc (M, 2N) fp32      (split epilogue)
    dt --dt_proj + bias + softplus--> delta (M, Di) act dtype (fp32 for fp32 activations or AUM_DELTA_16BIT=0)
    (u, delta, bc, z) --bidirectional scan--> out_z (M, Di)
    out_z --out_proj--> out (M, Dm)
No flip, no (b d l) transpose, no B/C rearrange copy is ever made.
"""
from __future__ import annotations

import os
import weakref
from typing import Optional

import torch

from . import _lib as L
from . import ops


_PREGATE_Z = os.environ.get("AUM_PREGATE_Z", "1") == "1"
# conv + SiLU fused into x_proj's operand producer (aum_conv_xproj_fwd, csrc/conv_xproj.cu): one launch and one pass over
# x instead of conv (write u) + x_proj (re-read u).  Measured at config 2: 0.082 ms per 64 sequences / 0.045 ms per 32
# against 0.053 + 0.035 / 0.031 + 0.025 ms for the two kernels it replaces (the first builds - 8 conv warps, CTA-wide
# barriers - were slower than the pair: profiles/r2_ncu_conv_xproj_v2_summary.txt).  AUM_FUSE_CONV_XPROJ=0 disables it.
_FUSE_CONV_XPROJ = os.environ.get("AUM_FUSE_CONV_XPROJ", "1") == "1"
# delta = softplus(dt_proj(dt) + bias) stored in the activation dtype when that is 16-bit (inference path): the reference
# rounds the dt_proj output to the autocast dtype as well (selective_scan_interface.py:468, before bias + softplus, where
# the rounding costs more), so the emulated-reference parity criterion (tests/test_parity_tiers_gpu.py) covers it.  It
# halves the largest tensor of the block (202 -> 101 MB written by dt_proj and read by the scan at config 2).
# AUM_DELTA_16BIT=0 keeps delta in fp32.
_DELTA_16BIT = os.environ.get("AUM_DELTA_16BIT", "1") == "1"


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


# Weight generation: bumped by anything that changes parameter VALUES behind autograd's back (the fused Adam kernel
# writes through raw pointers).  Cache entries and CUDA graphs are keyed on it in addition to the version counters.
_GENERATION = 0


def bump_generation() -> int:
    global _GENERATION
    _GENERATION += 1
    return _GENERATION


def generation() -> int:
    return _GENERATION


# Weight-gradient GEMMs on a side stream (training with a flat gradient buffer, aum_b200.dist.FlatGradReducer): dW = dY^T X
# is off the critical path of backward - nothing reads it before the all-reduce / optimiser - so it is launched on a second
# stream, ordered after the kernels that produce its operands, and fills SMs the main stream's kernels leave idle (the
# backward scan's 1.73-wave tail, the small kernels of the dt_proj chain).  The operands are kept alive until the join;
# FlatGradReducer joins before it launches an all-reduce piece and before the optimiser step.  AUM_WGRAD_SIDE=0 turns it off.
_WGRAD_SIDE = os.environ.get("AUM_WGRAD_SIDE", "1") == "1"


class _SideWork:
    def __init__(self):
        self.streams = {}
        self.pending = []
        self.dirty = set()
        self.cb_queued = False

    def launch(self, fn, *keep):
        dev = keep[0].device
        cur = torch.cuda.current_stream(dev)
        side = self.streams.get(dev)
        if side is None:
            side = self.streams[dev] = torch.cuda.Stream(device=dev)
        side.wait_stream(cur)                  # the operands are final
        with torch.cuda.stream(side):
            fn()
        self.pending.extend(keep)              # (their memory must not be handed out again before the join)
        self.dirty.add(dev)
        if not self.cb_queued:                 # whoever reads .grad after loss.backward() returns sees finished gradients
            from torch.autograd import Variable
            Variable._execution_engine.queue_callback(self._backward_done)
            self.cb_queued = True

    def _backward_done(self):
        self.cb_queued = False
        self.join()

    def join(self):
        for dev in self.dirty:
            torch.cuda.current_stream(dev).wait_stream(self.streams[dev])
        self.dirty.clear()
        self.pending.clear()


side_work = _SideWork()


class _DerivedCache:
    """Derived, read-only views of parameters (16-bit copies, zero-padded copies, A = -exp(A_log)).
    Entries are revalidated against the parameter's version counter and storage pointer, so in-place
    optimizer updates or load_state_dict invalidate them."""

    def __init__(self):
        self._d = {}

    def get(self, param: torch.Tensor, tag: str, fn):
        key = (id(param), tag)
        ver = (param._version, param.data_ptr(), param.device, param.dtype, _GENERATION)
        hit = self._d.get(key)
        if hit is not None and hit[0] == ver and hit[2]() is param:
            return hit[1]
        with torch.no_grad():
            val = fn(param.detach())
        if len(self._d) > 4096:      # ad-hoc (non-parameter) tensors: drop entries whose source is gone
            self._d = {k: v for k, v in self._d.items() if v[2]() is not None}
            if len(self._d) > 4096:
                self._d.clear()
        self._d[key] = (ver, val, weakref.ref(param))
        return val

    def clear(self):
        self._d.clear()


_cache = _DerivedCache()


def shadow16(param: torch.Tensor, dtype: torch.dtype) -> Optional[torch.Tensor]:
    """The optimiser-maintained 16-bit copy of `param` (aum_b200.dist.FlatAdam(shadow_dtype=...): written by the fused
    Adam kernel in the same pass as the fp32 update), or None when there is none, it has another dtype, or the parameter
    was modified behind the optimiser's back since (version counter moved)."""
    sh = getattr(param, "_aum_w16", None)
    if sh is None or sh.dtype != dtype or getattr(param, "_aum_w16_ver", None) != param._version:
        return None
    return sh


def _w(param: torch.Tensor, dtype: torch.dtype, pad_cols: Optional[int] = None) -> torch.Tensor:
    """Weight as contiguous `dtype`, optionally zero-padded to pad_cols columns."""
    if param.dim() == 2 and (pad_cols is None or pad_cols == param.shape[1]):
        sh = shadow16(param, dtype)
        if sh is not None:
            return sh                 # no cast kernel: the optimiser step already produced it

    def make(p):
        t = p.to(dtype)
        if pad_cols is not None and pad_cols != t.shape[1]:
            t2 = torch.zeros((t.shape[0], pad_cols), device=t.device, dtype=dtype)
            t2[:, : t.shape[1]] = t
            t = t2
        return t.contiguous()
    return _cache.get(param, f"w:{dtype}:{pad_cols}", make)


def _w2d(param: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """Weight of any rank as a contiguous (out, rest) `dtype` matrix (Linear weights, the patch-embedding conv)."""
    sh = shadow16(param, dtype)
    if sh is not None:
        return sh.reshape(sh.shape[0], -1)
    return _cache.get(param, f"w2d:{dtype}", lambda p: p.reshape(p.shape[0], -1).to(dtype).contiguous())


def _wT(param: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """W^T as a contiguous `dtype` matrix (the K-contiguous operand of dX = dY @ W), cached per parameter version: ONE
    transpose + cast launch (aum_transpose), reading the 16-bit shadow copy when there is one."""
    def make(p):
        sh = shadow16(param, dtype)
        src = (sh if sh is not None else p).reshape(p.shape[0], -1)
        return ops.transpose(src.unsqueeze(0), dst_dtype=dtype).squeeze(0)
    return _cache.get(param, f"wT:{dtype}", make)


def _f32(param: torch.Tensor) -> torch.Tensor:
    return _cache.get(param, "f32", lambda p: p.float().contiguous())


def _neg_exp(a_log: torch.Tensor) -> torch.Tensor:
    # A = -exp(A_log.float())   (mamba_simple.py:193,197)
    return _cache.get(a_log, "negexp", lambda p: (-torch.exp(p.float())).contiguous())


def _conv_w(weight: torch.Tensor) -> torch.Tensor:
    # "d 1 w -> d w"  (selective_scan_interface.py:460)
    return _cache.get(weight, "convw", lambda p: p.float().reshape(p.shape[0], p.shape[-1]).contiguous())


def _pipeline(xz: torch.Tensor, Di: int, N: int, conv_w, conv_b, x_proj_w, dt_proj_w, dt_bias,
              *, reverse: bool, delta_dtype: torch.dtype, backend: int):
    """conv -> x_proj -> dt_proj for one parameter set (selective_scan_interface.py:461-496).
    xz: (B, L, 2Di) token-major.  Returns u (B,L,Di), delta (B,L,Di), Bm, Cm (B,L,N) views."""
    B, Lq, _ = xz.shape
    M = B * Lq
    act = xz.dtype
    R = dt_proj_w.shape[1]
    x = xz[..., :Di]
    Rpad = _round_up(R, 8)
    dt = torch.empty((M, Rpad), device=xz.device, dtype=act)
    bc = torch.empty((M, 2 * N), device=xz.device, dtype=torch.float32)
    cw, cb, wx = _conv_w(conv_w), (_f32(conv_b) if conv_b is not None else None), _w(x_proj_w, act)
    if _FUSE_CONV_XPROJ and backend != L.GEMM_SIMT and ops.conv_xproj_eligible(x, cw, wx, R, 2 * N):
        # conv + SiLU as the producer of x_proj's tensor-core operand: one launch, x read once (:463 + :467)
        u = ops.conv_xproj(x, cw, cb, wx, R, dt, bc, reverse=reverse)
    else:
        u = ops.causal_conv1d(x, cw, cb, silu=True, reverse=reverse)
        ops.gemm_tn(u.view(M, Di), wx, out=dt, out2=bc, split=R, backend=backend)
    delta = ops.gemm_tn(dt, _w(dt_proj_w, act, pad_cols=Rpad), k=R, bias=_f32(dt_bias), act=L.ACT_SOFTPLUS,
                        out_dtype=delta_dtype, backend=backend)
    bc3 = bc.view(B, Lq, 2 * N)
    return u, delta.view(B, Lq, Di), bc3[..., :N], bc3[..., N:]


def mamba_mixer_forward(m, hidden: torch.Tensor, *, backend: int = L.GEMM_AUTO,
                        delta_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """Forward of one mixer.  ``m`` carries the reference module's parameters/attributes
    (in_proj, conv1d, x_proj, dt_proj, A_log, D, out_proj [, A_b_log, conv1d_b, x_proj_b, dt_proj_b, D_b, gamma]).
    hidden: (B, L, Dm) in the activation dtype (fp32 / fp16 / bf16).  Returns (B, L, Dm)."""
    L.require_cuda(hidden)
    B, Lq, Dm = hidden.shape
    M = B * Lq
    act = hidden.dtype
    Di, N = m.d_inner, m.d_state
    if delta_dtype is None:          # policy default: the activation dtype for 16-bit activations, else fp32
        delta_dtype = act if (_DELTA_16BIT and act != torch.float32) else torch.float32
    elif delta_dtype != torch.float32:
        delta_dtype = act
    h2 = hidden.reshape(M, Dm)
    if h2.stride(-1) != 1:
        h2 = h2.contiguous()
    in_b = _f32(m.in_proj.bias) if m.in_proj.bias is not None else None
    # in_proj; with pre-gating its epilogue already applies SiLU to the z half (columns >= Di), so that the
    # MUFU-bound scan only multiplies
    pregate = _PREGATE_Z
    xz = ops.gemm_tn(h2, _w(m.in_proj.weight, act), bias=in_b, backend=backend,
                     act=L.act_from(L.ACT_SILU, Di) if pregate else L.ACT_NONE).view(B, Lq, 2 * Di)    # (:185-191)
    z = xz[..., Di:]
    A = _neg_exp(m.A_log)
    Dv = _f32(m.D)
    kw = dict(delta_dtype=delta_dtype, backend=backend)
    bt = m.bimamba_type
    scale = 1.0
    if bt == "v1":      # Fo-Bi: shared projections, A vs A_b  (:198-213 -> BiMambaInnerFn.forward :441-517)
        u, delta, Bm, Cm = _pipeline(xz, Di, N, m.conv1d.weight, m.conv1d.bias, m.x_proj.weight,
                                     m.dt_proj.weight, m.dt_proj.bias, reverse=False, **kw)
        fwd = ops.ScanDirection(u, delta, A, Bm, Cm, Dv)
        bwd = ops.ScanDirection(u, delta, _neg_exp(m.A_b_log), Bm, Cm, Dv)
    elif bt == "v2":    # Bi-Bi: two full parameter sets, second on the reversed sequence (:214-246)
        u, delta, Bm, Cm = _pipeline(xz, Di, N, m.conv1d.weight, m.conv1d.bias, m.x_proj.weight,
                                     m.dt_proj.weight, m.dt_proj.bias, reverse=False, **kw)
        ub, deltab, Bb, Cb = _pipeline(xz, Di, N, m.conv1d_b.weight, m.conv1d_b.bias, m.x_proj_b.weight,
                                       m.dt_proj_b.weight, m.dt_proj_b.bias, reverse=True, **kw)
        fwd = ops.ScanDirection(u, delta, A, Bm, Cm, Dv)
        bwd = ops.ScanDirection(ub, deltab, _neg_exp(m.A_b_log), Bb, Cb, _f32(m.D_b))
        if m.if_devide_out:
            scale = 0.5                                                                    # (:246)
    elif bt == "none":  # Fo-Fo (:248-263 -> MambaInnerFn.forward :296-365)
        u, delta, Bm, Cm = _pipeline(xz, Di, N, m.conv1d.weight, m.conv1d.bias, m.x_proj.weight,
                                     m.dt_proj.weight, m.dt_proj.bias, reverse=False, **kw)
        fwd, bwd = ops.ScanDirection(u, delta, A, Bm, Cm, Dv), None
    else:
        raise ValueError(f"unknown bimamba_type {bt!r}")
    out_z = ops.selective_scan(fwd, bwd, z, out_scale=scale, z_pregated=pregate)
    out_b = _f32(m.out_proj.bias) if m.out_proj.bias is not None else None
    out = ops.gemm_tn(out_z.view(M, Di), _w(m.out_proj.weight, act), bias=out_b, backend=backend).view(B, Lq, Dm)  # (:517)
    gamma = getattr(m, "gamma", None)
    if getattr(m, "init_layer_scale", None) is not None and gamma is not None:
        out = out * gamma.to(out.dtype)                                                    # (:309-310)
    return out
