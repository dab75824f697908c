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

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev, dt = self.params[0].device, torch.float32
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, device=dev, dtype=dt)
        off = 0
        for p in self.params:
            if p.dtype != dt:
                raise ValueError("FlatGradReducer expects fp32 parameters (the reference keeps fp32 master weights)")
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def reduce(self):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(dist.get_world_size())
        return self.flat
