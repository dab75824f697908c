"""Host-side multi-GPU plumbing of the path.

The path shards by sequence: one process per GPU, weights replicated, no collective on the forward path.  The
training path has exactly one collective per step — the gradient all-reduce the reference gets from DDP inside
``accelerator.backward`` (/root/reference/src/traintest.py:39,168): here a single NCCL all-reduce (mean) of ONE flat
fp32 buffer holding every gradient (92.1 M x 4 B = 368 MB for AuM-Base), instead of DDP's bucket sequence.
Everything here is backend-agnostic (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import os
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None, device: Optional[torch.device] = None) -> int:
    """Initialise torch.distributed from torchrun's RANK/WORLD_SIZE/MASTER_* environment; returns the world size."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        kw = {}
        if device is not None and device.type == "cuda":
            kw["device_id"] = device
        dist.init_process_group(backend or ("nccl" if torch.cuda.is_available() else "gloo"), **kw)
    return world


def max_over_ranks(value: float, device) -> float:
    """Device-timed durations are reduced with MAX over ranks (bench.py's timing rule)."""
    t = torch.tensor([float(value)], device=device, dtype=torch.float64)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous [lo, hi) slice of n_items owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class FlatGradReducer:
    """The one collective of the training path: an all-reduce (mean) of ONE flat fp32 gradient buffer.

    Gradients are views into the flat buffer (set once) and the parameters are marked ``_aum_direct_grad``: the engine's
    backward kernels accumulate straight into the views (aum_b200.autograd), so the collective needs no packing copy
    and the backward no AccumulateGrad pass.

    Overlap with backward (what DDP's buckets do for the reference, /root/reference/src/traintest.py:39,168): pass the
    parameters in the order their gradients become final (`AudioMamba.grad_ready_order()`: head first, layer 0 and the
    embeddings last) together with `chunk_ends`, the element offsets at which a chunk of that order is complete;
    `hook(k)` returns a tensor hook that launches chunk k's all-reduce on NCCL's stream the moment backward reaches
    that point, and `reduce()` launches whatever is left and joins.  Without chunk information `reduce()` is a single
    all-reduce after backward."""

    ALIGN = 8      # elements

    def __init__(self, params: Iterable[torch.nn.Parameter], chunk_after: Optional[List[int]] = None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev, dt = self.params[0].device, torch.float32
        # every tensor starts on a 32-byte boundary of the flat buffer (kernels take 16-byte vector accesses on
        # weights and weight gradients); the padding elements stay zero
        self.offsets, off = [], 0
        for p in self.params:
            if p.dtype != dt:
                raise ValueError("FlatGradReducer expects fp32 parameters (the reference keeps fp32 master weights)")
            self.offsets.append(off)
            off += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.numel = off
        self.flat = torch.zeros(self.numel, device=dev, dtype=dt)
        for p, o in zip(self.params, self.offsets):
            p.grad = self.flat[o:o + p.numel()].view_as(p)
            p._aum_direct_grad = True
        # chunk k = flat[bounds[k] : bounds[k+1]]; chunk_after lists, per chunk but the last, the number of leading
        # parameters (of the given order) it ends after
        ends = [self.offsets[i] if i < len(self.offsets) else self.numel for i in (chunk_after or [])]
        self.bounds = [0] + [e for e in ends if 0 < e < self.numel] + [self.numel]
        self._launched = 0
        self._works = []
        self.async_launches = 0

    @property
    def n_chunks(self) -> int:
        return len(self.bounds) - 1

    def zero(self):
        self.flat.zero_()
        self._launched = 0
        self._works = []

    def _active(self) -> bool:
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    def _launch_upto(self, k_end: int):
        """All-reduce chunks [launched, k_end) as ONE contiguous slice, asynchronously (NCCL orders it after the
        kernels already enqueued on the current stream)."""
        if k_end <= self._launched:
            return
        from . import mixer
        mixer.side_work.join()          # weight-gradient GEMMs launched on the side stream accumulate into this buffer
        lo, hi = self.bounds[self._launched], self.bounds[k_end]
        self._launched = k_end
        if not self._active() or hi <= lo:
            return
        sl = self.flat[lo:hi]
        if dist.get_backend() == "nccl":
            self._works.append((dist.all_reduce(sl, op=dist.ReduceOp.AVG, async_op=True), None))
        else:       # gloo (CPU tests) has no AVG
            self._works.append((dist.all_reduce(sl, op=dist.ReduceOp.SUM, async_op=True), sl))
        self.async_launches += 1

    def hook(self, k: int):
        """Tensor hook for the activation whose gradient is computed right after every parameter gradient of chunks
        0..k is final: launches those chunks' all-reduce while backward continues."""
        def fire(grad):
            self._launch_upto(k + 1)
            return grad
        return fire

    def reduce(self):
        """Launch what has not been launched yet, wait for every piece (the current stream waits; the host does not
        block on NCCL), and return the averaged flat gradient."""
        self._launch_upto(self.n_chunks)
        from . import mixer
        mixer.side_work.join()
        world = dist.get_world_size() if self._active() else 1
        for w, sl in self._works:
            w.wait()
            if sl is not None:
                sl.div_(world)
        self._works = []
        return self.flat


class FlatAdam:
    """torch.optim.Adam over ONE flat parameter buffer, one fused kernel per step.

    Parameters become views into a flat fp32 buffer (same order as the FlatGradReducer's gradient buffer), the two
    moments are flat as well, and `step()` is a single aum_adam_step launch instead of the multi-tensor
    implementation's ~10 passes.  Same update rule and defaults as the reference's recipe
    (/root/reference/src/traintest.py:32-34).  CUDA only (the product path has no CPU fallback).

    Mixed precision: bf16 and fp32 activations need no loss scaling.  With fp16 activations pass `grad_scale` = 1 / loss
    scale and `skip_nonfinite=True`: like accelerate's GradScaler (which the reference's fp16 recipe relies on) the step
    is skipped - p, m and v untouched - when the flat gradient holds an inf or NaN, and `step()` returns False so the
    caller can lower its scale."""

    def __init__(self, reducer: FlatGradReducer, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, shadow_dtype: Optional[torch.dtype] = None):
        self.reducer, self.lr, self.betas, self.eps, self.weight_decay = reducer, lr, tuple(betas), eps, weight_decay
        self.t = 0
        self.generation = 0          # bumped by every applied step: derived-weight caches / CUDA graphs key on versions
        flat = torch.zeros_like(reducer.flat)
        with torch.no_grad():
            for p, off in zip(reducer.params, reducer.offsets):
                n = p.numel()
                flat[off:off + n].copy_(p.detach().reshape(-1))
                p.data = flat[off:off + n].view_as(p)
        self.flat_p = flat
        self.m = torch.zeros_like(flat)
        self.v = torch.zeros_like(flat)
        # the step number also lives on the device (aum_adam_step_dev increments it and forms the bias corrections from
        # it): no argument of the launch changes from step to step, which is what lets TrainStep replay a CUDA graph
        self.step_dev = torch.zeros((), device=flat.device, dtype=torch.int32)
        # 16-bit shadow of the parameters, written by the Adam kernel in the same pass (shadow_dtype = the activation
        # dtype of the model): mixer._w / _w2d hand out views of it instead of launching one cast per weight per step
        self.flat16 = None
        if shadow_dtype is not None and shadow_dtype != torch.float32:
            self.flat16 = flat.to(shadow_dtype)
            for p, off in zip(reducer.params, reducer.offsets):
                p._aum_w16 = self.flat16[off:off + p.numel()].view_as(p)
            self._mark_shadow_current()

    def _mark_shadow_current(self):
        if self.flat16 is not None:
            for p in self.reducer.params:
                p._aum_w16_ver = p._version

    def resync_shadow(self):
        """Call after changing parameter values by anything but step() (load_state_dict, manual edits)."""
        if self.flat16 is not None:
            self.flat16.copy_(self.flat_p)
            self._mark_shadow_current()

    def after_device_step(self):
        """Host-side bookkeeping of one applied update (also called by TrainStep after a CUDA-graph replay, where no
        Python ran): step count, weight generation (every derived-weight cache entry - transposed copies, -exp(A_log),
        ... - and every captured inference graph is keyed on it, so both are rebuilt before the next eager forward)
        and, where torch exposes it, the parameters' version counters for third-party observers."""
        from . import mixer
        self.t += 1
        mixer.bump_generation()
        bump = getattr(torch._C, "_increment_version", None)
        if bump is not None:
            try:
                bump(self.reducer.params)
            except (TypeError, RuntimeError):
                try:
                    for p in self.reducer.params:
                        bump(p)
                except (TypeError, RuntimeError):
                    pass
        self._mark_shadow_current()
        self.generation += 1

    def step(self, grad_scale: float = 1.0, skip_nonfinite: bool = False) -> bool:
        from . import ops
        if skip_nonfinite and not bool(torch.isfinite(self.reducer.flat).all()):     # (one host sync, fp16 recipe only)
            return False
        ops.adam_step_dev(self.flat_p, self.reducer.flat, self.m, self.v, self.step_dev, lr=self.lr, betas=self.betas,
                          eps=self.eps, weight_decay=self.weight_decay, grad_scale=grad_scale, p16=self.flat16)
        # the kernel wrote through raw pointers, which autograd's version counters do not see
        self.after_device_step()
        return True

    def zero_grad(self):
        self.reducer.zero()
