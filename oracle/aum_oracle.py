"""CPU oracle for the AuM bidirectional selective-scan hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``audio-mamba-aum_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or as
the CPU baseline, never as the thing that is shipped.

This file is an independent *restatement* (plain torch-on-CPU, fp32 or fp64) of
the reference's own Python reference implementations.  Each function cites the
reference lines it follows (paths relative to /root/reference).  Parity pin:
``tests/golden/*.pt`` were produced by importing the real reference in the build
container (``oracle/make_golden.py`` + ``oracle/ref_loader.py``) and
``tests/test_oracle_cpu.py`` checks this restatement against them, so the oracle
is pinned to the reference's ``selective_scan_ref`` / ``bimamba_inner_ref`` /
``mamba_inner_ref`` / ``rms_norm_ref`` / ``AudioMamba.forward``.

The CUDA kernels of the pinned pip wheels (selective_scan_cuda, causal_conv1d_cuda;
mamba_ssm==1.1.3.post1, causal_conv1d==1.1.3.post1, README.md:58) are not in the
reference tree and have no golden vectors there: parity at *that* boundary is
"unpinned"; parity is pinned to the in-repo Python ``*_ref`` functions, which is
what BASELINE.json's north_star names.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# elementary pieces
# --------------------------------------------------------------------------------------

def softplus_oracle(x: torch.Tensor) -> torch.Tensor:
    """torch softplus, beta=1, threshold=20 (selective_scan_interface.py:106-107)."""
    return torch.where(x > 20.0, x, torch.log1p(torch.exp(torch.clamp(x, max=20.0))))


def silu_oracle(x: torch.Tensor) -> torch.Tensor:
    return x * torch.sigmoid(x)


def _rnd(t: Optional[torch.Tensor], io_dtype: Optional[torch.dtype]) -> Optional[torch.Tensor]:
    """Round-trip through the autocast dtype: the reference's half-precision rounding points.
    Under ``--mixed_precision=fp16`` (exps/*/*.sh) the reference materialises xz, conv1d_out, x_dbl (hence the raw
    delta, B and C), both directional out_z tensors, their sum and the out_proj output in the autocast dtype and
    casts the three projection weights to it (selective_scan_interface.py:452-457,463-468,499-507,517); scan
    state, A, D and delta_bias stay fp32.  ``io_dtype=None`` (default) leaves everything in fp32 (the oracle)."""
    if t is None or io_dtype is None:
        return t
    return t.to(io_dtype).to(t.dtype)


def causal_conv1d_oracle(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor],
                         silu: bool = True) -> torch.Tensor:
    """Depthwise causal conv along the last axis + optional SiLU.

    x: (B, D, L); weight: (D, W); bias: (D,) or None.
    Follows the reference's own fallback expression
    ``act(conv1d(x)[..., :seqlen])`` with ``padding=d_conv-1``
    (vim-mamba_ssm/mamba_ssm/modules/mamba_simple.py:272 and :76-84), written out as
    the explicit tap sum  out[b,d,l] = bias[d] + sum_k w[d,k] * x[b,d,l-(W-1)+k].
    """
    Bsz, D, L = x.shape
    W = weight.shape[-1]
    acc_dtype = torch.float64 if x.dtype == torch.float64 else torch.float32
    xf = x.to(acc_dtype)
    wf = weight.to(acc_dtype)
    xp = F.pad(xf, (W - 1, 0))
    out = torch.zeros(Bsz, D, L, dtype=acc_dtype)
    for k in range(W):
        out = out + wf[:, k].view(1, D, 1) * xp[:, :, k:k + L]
    if bias is not None:
        out = out + bias.to(acc_dtype).view(1, D, 1)
    if silu:
        out = silu_oracle(out)
    return out.to(x.dtype)


def selective_scan_oracle(u, delta, A, B, C, D=None, z=None, delta_bias=None,
                          delta_softplus=False, return_last_state=False,
                          compute_dtype=torch.float32):
    """One scan direction.  Restates selective_scan_ref
    (vim-mamba_ssm/mamba_ssm/ops/selective_scan_interface.py:86-152), real A only.

    u, delta, z: (B, D, L);  A: (D, N);  B, C: (B, N, L) or (B, G, N, L);
    D, delta_bias: (D,).  Internals in ``compute_dtype`` (the reference hard-casts to
    fp32 at :102-103,:117-118; fp64 is offered to bound the oracle's own rounding).
    The recurrence (:133-146):  h_l = exp(delta_l*A) * h_{l-1} + delta_l*B_l*u_l ;
    y_l = <C_l, h_l> ;  out = (y + D*u) * silu(z)  (:148-150), cast to u.dtype (:151).
    """
    dtype_in = u.dtype
    cd = compute_dtype
    u_ = u.to(cd)
    delta_ = delta.to(cd)
    if delta_bias is not None:
        delta_ = delta_ + delta_bias.to(cd)[..., None]
    if delta_softplus:
        delta_ = softplus_oracle(delta_)
    Bsz, Dm, L = u_.shape
    N = A.shape[1]
    A_ = A.to(cd)
    B_ = B.to(cd)
    C_ = C.to(cd)
    if B_.dim() == 4:   # (B, G, N, L) -> (B, D, N, L)   (:128)
        B_ = B_.repeat_interleave(Dm // B_.shape[1], dim=1)
    if C_.dim() == 4:   # (:131)
        C_ = C_.repeat_interleave(Dm // C_.shape[1], dim=1)
    h = torch.zeros(Bsz, Dm, N, dtype=cd)
    ys = torch.empty(Bsz, Dm, L, dtype=cd)
    for i in range(L):
        dA = torch.exp(delta_[:, :, i, None] * A_[None])                 # (B, D, N)   (:121)
        if B_.dim() == 3:
            dBu = delta_[:, :, i, None] * B_[:, None, :, i] * u_[:, :, i, None]   # (:126)
        else:
            dBu = delta_[:, :, i, None] * B_[:, :, :, i] * u_[:, :, i, None]      # (:129)
        h = dA * h + dBu                                                  # (:134)
        if C_.dim() == 3:
            ys[:, :, i] = (h * C_[:, None, :, i]).sum(-1)                 # (:139)
        else:
            ys[:, :, i] = (h * C_[:, :, :, i]).sum(-1)                    # (:141)
    out = ys if D is None else ys + u_ * D.to(cd)[None, :, None]          # (:148)
    if z is not None:
        out = out * silu_oracle(z.to(cd))                                 # (:150)
    out = out.to(dtype_in)
    return (out, h) if return_last_state else out


# --------------------------------------------------------------------------------------
# fused inner ops
# --------------------------------------------------------------------------------------

def _proj_stage(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, d_state, io_dtype=None):
    """conv -> x_proj -> dt_proj, shared by the three inner ops
    (selective_scan_interface.py:642-668 / :679-705).  io_dtype: see _rnd (emulated-reference rounding points
    of the fused forward, :452-468)."""
    L = xz.shape[-1]
    R = delta_proj_weight.shape[1]
    xz = _rnd(xz, io_dtype)
    x, z = xz.chunk(2, dim=1)
    w = conv1d_weight.reshape(conv1d_weight.shape[0], conv1d_weight.shape[-1])   # "d 1 w -> d w"
    x = _rnd(causal_conv1d_oracle(x, w, conv1d_bias, silu=True), io_dtype)       # (:646 / :683; :463)
    Bsz, Di, _ = x.shape
    x_dbl = _rnd(F.linear(x.permute(0, 2, 1).reshape(Bsz * L, Di), _rnd(x_proj_weight, io_dtype)), io_dtype)  # (:650; :467)
    delta = _rnd((_rnd(delta_proj_weight, io_dtype) @ x_dbl[:, :R].t()), io_dtype).reshape(Di, Bsz, L).permute(1, 0, 2)   # (:651-652; :468)
    Bm = x_dbl[:, R:R + d_state].reshape(Bsz, L, d_state).permute(0, 2, 1).contiguous()   # (:654-658)
    Cm = x_dbl[:, -d_state:].reshape(Bsz, L, d_state).permute(0, 2, 1).contiguous()       # (:662-666)
    return x, z, delta, Bm, Cm


def mamba_inner_oracle(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight,
                       out_proj_weight, out_proj_bias, A, D=None, delta_bias=None,
                       compute_dtype=torch.float32, io_dtype=None):
    """Fo-Fo inner op; restates mamba_inner_ref (selective_scan_interface.py:636-670)."""
    x, z, delta, Bm, Cm = _proj_stage(xz, conv1d_weight, conv1d_bias, x_proj_weight,
                                      delta_proj_weight, A.shape[-1], io_dtype)
    y = _rnd(selective_scan_oracle(x, delta, A, Bm, Cm, D, z=z, delta_bias=delta_bias,
                                   delta_softplus=True, compute_dtype=compute_dtype), io_dtype)    # (:669)
    return _rnd(F.linear(y.permute(0, 2, 1), _rnd(out_proj_weight, io_dtype), _rnd(out_proj_bias, io_dtype)), io_dtype)   # (:670)


def mamba_inner_no_out_proj_oracle(xz, conv1d_weight, conv1d_bias, x_proj_weight,
                                   delta_proj_weight, A, D=None, delta_bias=None,
                                   compute_dtype=torch.float32, io_dtype=None):
    """One Bi-Bi pipeline: MambaInnerFnNoOutProj.forward semantics
    (selective_scan_interface.py:155-224) = mamba_inner_ref without :670; returns (B, Di, L)."""
    x, z, delta, Bm, Cm = _proj_stage(xz, conv1d_weight, conv1d_bias, x_proj_weight,
                                      delta_proj_weight, A.shape[-1], io_dtype)
    return _rnd(selective_scan_oracle(x, delta, A, Bm, Cm, D, z=z, delta_bias=delta_bias,
                                      delta_softplus=True, compute_dtype=compute_dtype), io_dtype)


def bimamba_inner_oracle(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight,
                         out_proj_weight, out_proj_bias, A, A_b, D=None, delta_bias=None,
                         compute_dtype=torch.float32, io_dtype=None):
    """Fo-Bi inner op; restates bimamba_inner_ref (selective_scan_interface.py:673-709).

    Both directions share conv / x_proj / dt_proj / D / delta_bias; only A vs A_b
    differ; the reverse scan runs on flipped tensors and is flipped back (:706-708).
    D*u therefore enters twice."""
    x, z, delta, Bm, Cm = _proj_stage(xz, conv1d_weight, conv1d_bias, x_proj_weight,
                                      delta_proj_weight, A.shape[-1], io_dtype)
    y = _rnd(selective_scan_oracle(x, delta, A, Bm, Cm, D, z=z, delta_bias=delta_bias,
                                   delta_softplus=True, compute_dtype=compute_dtype), io_dtype)         # (:706; out_z_f :499)
    y_b = _rnd(selective_scan_oracle(x.flip([-1]), delta.flip([-1]), A_b, Bm.flip([-1]),
                                     Cm.flip([-1]), D, z.flip([-1]), delta_bias,
                                     delta_softplus=True, compute_dtype=compute_dtype), io_dtype)       # (:707; out_z_b :503)
    y = _rnd(y + y_b.flip([-1]), io_dtype)                                              # (:708; :507)
    return _rnd(F.linear(y.permute(0, 2, 1), _rnd(out_proj_weight, io_dtype), _rnd(out_proj_bias, io_dtype)), io_dtype)   # (:709; :517)


# --------------------------------------------------------------------------------------
# Mamba mixer module forward (functional, from a parameter dict)
# --------------------------------------------------------------------------------------

def mamba_forward_oracle(p: Dict[str, torch.Tensor], hidden: torch.Tensor, bimamba_type: str = "v1",
                         if_devide_out: bool = False, compute_dtype=torch.float32, io_dtype=None) -> torch.Tensor:
    """Mamba.forward fast path (vim-mamba_ssm/mamba_ssm/modules/mamba_simple.py:169-311).

    ``p`` uses the module's state-dict key names (in_proj.weight, conv1d.weight, ...).
    hidden: (B, L, Dm) -> (B, L, Dm).  io_dtype: emulate the reference's autocast rounding points (see _rnd);
    the in_proj matmul (:185-189) runs in the autocast dtype, so its operands and result are rounded too."""
    Bsz, L, Dm = hidden.shape
    W_in = _rnd(p["in_proj.weight"], io_dtype)
    hidden = _rnd(hidden, io_dtype)
    xz = (W_in @ hidden.reshape(Bsz * L, Dm).t()).reshape(-1, Bsz, L).permute(1, 0, 2)   # (:185-189)
    if "in_proj.bias" in p and p["in_proj.bias"] is not None:
        xz = xz + p["in_proj.bias"].to(xz.dtype)[None, :, None]                          # (:190-191)
    A = -torch.exp(p["A_log"].float())                                                   # (:193)
    ob = p.get("out_proj.bias", None)
    if bimamba_type == "v1":
        A_b = -torch.exp(p["A_b_log"].float())                                           # (:197)
        out = bimamba_inner_oracle(xz, p["conv1d.weight"], p["conv1d.bias"], p["x_proj.weight"],
                                   p["dt_proj.weight"], p["out_proj.weight"], ob, A, A_b,
                                   p["D"].float(), p["dt_proj.bias"].float(), compute_dtype, io_dtype)   # (:198-213)
    elif bimamba_type == "v2":
        A_b = -torch.exp(p["A_b_log"].float())                                           # (:215)
        o = mamba_inner_no_out_proj_oracle(xz, p["conv1d.weight"], p["conv1d.bias"],
                                           p["x_proj.weight"], p["dt_proj.weight"], A,
                                           p["D"].float(), p["dt_proj.bias"].float(), compute_dtype, io_dtype)  # (:216-228)
        o_b = mamba_inner_no_out_proj_oracle(xz.flip([-1]), p["conv1d_b.weight"], p["conv1d_b.bias"],
                                             p["x_proj_b.weight"], p["dt_proj_b.weight"], A_b,
                                             p["D_b"].float(), p["dt_proj_b.bias"].float(), compute_dtype, io_dtype)  # (:229-241)
        y = _rnd((o + o_b.flip([-1])).permute(0, 2, 1), io_dtype)
        if if_devide_out:
            y = y / 2                                                                    # (:246)  (exact in fp16/bf16)
        out = _rnd(F.linear(y, _rnd(p["out_proj.weight"], io_dtype), _rnd(ob, io_dtype)), io_dtype)   # (:244/:246)
    elif bimamba_type == "none":
        out = mamba_inner_oracle(xz, p["conv1d.weight"], p["conv1d.bias"], p["x_proj.weight"],
                                 p["dt_proj.weight"], p["out_proj.weight"], ob, A,
                                 p["D"].float(), p["dt_proj.bias"].float(), compute_dtype, io_dtype)     # (:249-263)
    else:
        raise ValueError(bimamba_type)
    if "gamma" in p and p["gamma"] is not None:
        out = _rnd(out * p["gamma"], io_dtype)                                           # (:309-310)
    return out


# --------------------------------------------------------------------------------------
# caller-side ops (next rows of SURVEY.md section 8f) and whole-model forward
# --------------------------------------------------------------------------------------

def rms_norm_oracle(x, weight, bias=None, residual=None, eps=1e-5, prenorm=False):
    """rms_norm_ref with upcast=True (vim-mamba_ssm/mamba_ssm/ops/triton/layernorm.py:35-48):
    add residual in fp32, rstd = 1/sqrt(mean(x^2)+eps), out = x*rstd*w (+b), cast to x.dtype;
    prenorm returns (out, x_plus_residual_fp32)."""
    dtype = x.dtype
    xf = x.float()
    if residual is not None:
        xf = xf + residual.float()
    rstd = 1.0 / torch.sqrt(xf.square().mean(dim=-1, keepdim=True) + eps)
    out = xf * rstd * weight.float()
    if bias is not None:
        out = out + bias.float()
    out = out.to(dtype)
    return (out, xf) if prenorm else out


def block_params(sd: Dict[str, torch.Tensor], i: int) -> Dict[str, torch.Tensor]:
    pre = f"layers.{i}.mixer."
    return {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}


def audio_mamba_forward_oracle(sd: Dict[str, torch.Tensor], x: torch.Tensor, *, depth: int,
                               bimamba_type: str = "v1", if_devide_out: bool = True,
                               patch: Tuple[int, int] = (16, 16), eps: float = 1e-5,
                               compute_dtype=torch.float32, n_blocks: Optional[int] = None,
                               return_features: bool = False, io_dtype=None) -> torch.Tensor:
    """AudioMamba.forward, default configuration (src/models/mamba_models.py:509-685):
    rms_norm=True, fused_add_norm=True, residual_in_fp32=True, abs pos-embed, middle cls token,
    no rope, no flips, drop_path 0.  ``sd`` is the model's state dict.  x: (B, T, F).

    n_blocks (< depth) runs only the first n_blocks layers — used ONLY by bench.py's bounded CPU
    sample, never by parity tests.

    io_dtype (torch.float16 / torch.bfloat16): the *emulated reference* — the same forward with every tensor
    the reference materialises in the autocast dtype under ``accelerate --mixed_precision`` rounded to it
    (patch-embed conv, the normed activations, everything listed at _rnd, the head); the residual stream stays
    fp32 (residual_in_fp32, mamba_models.py:209).  Used by the 16-bit parity tier: SURVEY.md section 8(c)."""
    Bsz = x.shape[0]
    img = _rnd(x.unsqueeze(1).transpose(2, 3), io_dtype)                    # (:510-511) B,1,F,T
    w = _rnd(sd["patch_embed.proj.weight"], io_dtype)
    b = _rnd(sd["patch_embed.proj.bias"], io_dtype)
    t = _rnd(F.conv2d(img, w, b, stride=patch), io_dtype)                   # tokenization.py:306
    t = t.flatten(2).transpose(1, 2)                                        # (:308) B, N, Dm
    N = t.shape[1]
    tp = N // 2                                                             # (:528-529)
    cls = sd["cls_token"].expand(Bsz, -1, -1)
    t = torch.cat((t[:, :tp], cls, t[:, tp:]), dim=1)                       # (:534)
    # pos-embed: slot 0 belongs to the cls token (tokenization.py:414-451: insert_to_prefix,
    # add, insert_from_prefix)
    pe = sd["pos_embed.pos_embed"]                                          # (1, N+1, Dm)
    pe_seq = torch.cat((pe[:, 1:tp + 1], pe[:, :1], pe[:, tp + 1:]), dim=1)
    hidden = t + pe_seq
    residual = None
    nb = depth if n_blocks is None else n_blocks
    for i in range(nb):                                                     # (:602-622)
        nw = sd[f"layers.{i}.norm.weight"]
        hidden, residual = rms_norm_oracle(hidden, nw, None, residual, eps, prenorm=True)   # (:77-97)
        hidden = mamba_forward_oracle(block_params(sd, i), _rnd(hidden, io_dtype), bimamba_type,
                                      if_devide_out, compute_dtype, io_dtype)                # (:98)
    hidden = _rnd(rms_norm_oracle(hidden, sd["norm_f.weight"], None, residual, eps, prenorm=False), io_dtype)  # (:646-657)
    feat = hidden[:, tp, :]                                                 # (:660-664)
    if return_features:
        return feat
    return _rnd(F.linear(feat, _rnd(sd["head.weight"], io_dtype), _rnd(sd["head.bias"], io_dtype)), io_dtype)   # (:682)


# --------------------------------------------------------------------------------------
# deterministic parameter / input generators shared by tests, smoke() and bench.py
# --------------------------------------------------------------------------------------

def make_mamba_params(d_model: int, *, d_state: int = 16, d_conv: int = 4, expand: int = 2,
                      bimamba_type: str = "v1", seed: int = 3949, perturb_A: float = 0.0,
                      dt_min: float = 1e-3, dt_max: float = 0.1) -> Dict[str, torch.Tensor]:
    """Parameters with the reference's shapes and init distributions
    (mamba_simple.py:74-167): nn.Linear/Conv1d default U(+-1/sqrt(fan_in)), dt init (:95-113),
    S4D-real A_log = log(1..N) (:116-123), D = 1 (:126).  perturb_A adds N(0, perturb_A^2) to
    A_log/A_b_log so a structured-A shortcut cannot be what is exercised."""
    g = torch.Generator().manual_seed(seed)
    Di = expand * d_model
    R = math.ceil(d_model / 16)

    def U(shape, fan_in):
        bound = 1.0 / math.sqrt(fan_in)
        return (torch.rand(shape, generator=g) * 2 - 1) * bound

    def dt_bias():
        dt = torch.exp(torch.rand(Di, generator=g) * (math.log(dt_max) - math.log(dt_min))
                       + math.log(dt_min)).clamp(min=1e-4)
        return dt + torch.log(-torch.expm1(-dt))

    def a_log():
        a = torch.log(torch.arange(1, d_state + 1, dtype=torch.float32)).repeat(Di, 1)
        if perturb_A > 0:
            a = a + perturb_A * torch.randn(Di, d_state, generator=g)
        return a.contiguous()

    p = {
        "in_proj.weight": U((2 * Di, d_model), d_model),
        "conv1d.weight": U((Di, 1, d_conv), d_conv),
        "conv1d.bias": U((Di,), d_conv),
        "x_proj.weight": U((R + 2 * d_state, Di), Di),
        "dt_proj.weight": U((Di, R), R),   # U(+-R^-0.5)  (:95-99)
        "dt_proj.bias": dt_bias(),
        "A_log": a_log(),
        "D": torch.ones(Di),
        "out_proj.weight": U((d_model, Di), Di),
    }
    if bimamba_type in ("v1", "v2"):
        p["A_b_log"] = a_log()
    if bimamba_type == "v2":
        p["conv1d_b.weight"] = U((Di, 1, d_conv), d_conv)
        p["conv1d_b.bias"] = U((Di,), d_conv)
        p["x_proj_b.weight"] = U((R + 2 * d_state, Di), Di)
        p["dt_proj_b.weight"] = U((Di, R), R)
        p["dt_proj_b.bias"] = dt_bias()
        p["D_b"] = torch.ones(Di)
    return p


def make_audio_mamba_state(embed_dim: int, depth: int, *, num_classes: int = 527,
                           spectrogram_size: Tuple[int, int] = (128, 1024),
                           patch: Tuple[int, int] = (16, 16), bimamba_type: str = "v1",
                           seed: int = 3949, perturb_A: float = 0.0) -> Dict[str, torch.Tensor]:
    """A full AudioMamba state dict (key names of src/models/mamba_models.py) with random-init
    weights of the reference's shapes; out_proj rescaled by 1/sqrt(depth) as _init_weights does
    (mamba_models.py:164-172)."""
    g = torch.Generator().manual_seed(seed + 1)
    F_, T_ = spectrogram_size
    n_patches = (F_ // patch[0]) * (T_ // patch[1])
    sd: Dict[str, torch.Tensor] = {}
    fan_in = patch[0] * patch[1]
    sd["patch_embed.proj.weight"] = torch.randn(embed_dim, 1, patch[0], patch[1], generator=g) / math.sqrt(fan_in)
    sd["patch_embed.proj.bias"] = torch.zeros(embed_dim)
    sd["cls_token"] = 0.02 * torch.randn(1, 1, embed_dim, generator=g)
    sd["pos_embed.pos_embed"] = 0.02 * torch.randn(1, n_patches + 1, embed_dim, generator=g)
    for i in range(depth):
        p = make_mamba_params(embed_dim, bimamba_type=bimamba_type, seed=seed + 17 * (i + 1),
                              perturb_A=perturb_A)
        p["out_proj.weight"] = p["out_proj.weight"] / math.sqrt(depth)
        for k, v in p.items():
            sd[f"layers.{i}.mixer.{k}"] = v
        sd[f"layers.{i}.norm.weight"] = torch.ones(embed_dim) + 0.05 * torch.randn(embed_dim, generator=g)
    sd["norm_f.weight"] = torch.ones(embed_dim) + 0.05 * torch.randn(embed_dim, generator=g)
    sd["head.weight"] = 0.02 * torch.randn(num_classes, embed_dim, generator=g)
    sd["head.bias"] = torch.zeros(num_classes)
    return sd


def make_spectrogram(batch: int, spectrogram_size: Tuple[int, int] = (128, 1024), seed: int = 3949):
    """(B, T, F) fp32 ~ N(0, 0.5^2): the dataloader normalises to about half-unit std
    (src/dataloader.py:221)."""
    g = torch.Generator().manual_seed(seed + 7)
    F_, T_ = spectrogram_size
    return 0.5 * torch.randn(batch, T_, F_, generator=g)
