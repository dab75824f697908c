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
    """One all-reduce per step over a single flat gradient buffer.

    Gradients are views into the flat buffer (set once), so backward writes land in it directly and the
    collective needs no packing copies.  `reduce()` averages over ranks."""

    ALIGN = 8      # elements

    def __init__(self, params: Iterable[torch.nn.Parameter]):
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

    def zero(self):
        self.flat.zero_()

    def reduce(self):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(dist.get_world_size())
        return self.flat


class FlatAdam:
    """torch.optim.Adam over ONE flat parameter buffer, one fused kernel per step.

    Parameters become views into a flat fp32 buffer (same order as the FlatGradReducer's gradient buffer), the two
    moments are flat as well, and `step()` is a single aum_adam_step launch instead of the multi-tensor
    implementation's ~10 passes.  Same update rule and defaults as the reference's recipe
    (/root/reference/src/traintest.py:32-34).  CUDA only (the product path has no CPU fallback)."""

    def __init__(self, reducer: FlatGradReducer, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0):
        self.reducer, self.lr, self.betas, self.eps, self.weight_decay = reducer, lr, tuple(betas), eps, weight_decay
        self.t = 0
        flat = torch.zeros_like(reducer.flat)
        with torch.no_grad():
            for p, off in zip(reducer.params, reducer.offsets):
                n = p.numel()
                flat[off:off + n].copy_(p.detach().reshape(-1))
                p.data = flat[off:off + n].view_as(p)
        self.flat_p = flat
        self.m = torch.zeros_like(flat)
        self.v = torch.zeros_like(flat)

    def step(self, grad_scale: float = 1.0):
        from . import mixer, ops
        self.t += 1
        ops.adam_step(self.flat_p, self.reducer.flat, self.m, self.v, lr=self.lr, betas=self.betas, eps=self.eps,
                      weight_decay=self.weight_decay, step=self.t, grad_scale=grad_scale)
        # the kernel wrote through a raw pointer: tell autograd's version counters (the derived 16-bit weight copies of
        # the mixer are revalidated against them)
        bump = getattr(torch._C, "_increment_version", None)
        if bump is not None:
            try:
                bump(self.reducer.params)
            except TypeError:
                for p in self.reducer.params:
                    bump(p)
        else:
            mixer._cache.clear()

    def zero_grad(self):
        self.reducer.zero()
