"""One data-parallel training step of AudioMamba on the B200 engine (BASELINE config 3).

Mirrors the body of the reference's loop (/root/reference/src/traintest.py:144-169): forward in the mixed-precision
dtype, BCE-with-logits (or cross-entropy) loss, zero_grad, backward, optimizer step - with the reference's DDP bucket
all-reduce (traintest.py:39,168) replaced by ONE flat fp32 gradient buffer whose all-reduce is launched in a few
reverse-layer-order chunks from autograd hooks while backward is still running, and torch.optim.Adam replaced by one
fused kernel over the flat parameter / moment buffers.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from .dist import FlatAdam, FlatGradReducer


class TrainStep:
    """model: aum_b200.audio_mamba.AudioMamba (fp32 parameters, act_dtype = the mixed-precision dtype).
    n_chunks: pieces the gradient all-reduce is split into (last layers first).  Adam defaults: the reference's recipe
    (traintest.py:32-34: betas (0.95, 0.999), weight_decay 5e-7, lr from the experiment script).
    shadow16: the fused Adam kernel also writes the parameters' copy in the activation dtype (no per-weight cast kernels
    in the next forward).  cuda_graph: capture the WHOLE step (zero, forward, loss, backward incl. the all-reduce pieces
    launched from its hooks, Adam with a device-side step counter) once per input shape and replay it; inputs are copied
    into the captured buffers, the returned loss is the captured tensor (overwritten by the next call)."""

    def __init__(self, model, lr: float = 1e-5, betas=(0.95, 0.999), eps: float = 1e-8, weight_decay: float = 5e-7,
                 n_chunks: int = 3, loss: str = "bce", shadow16: bool = True, cuda_graph: bool = False,
                 loss_scale: float = 1.0):
        self.model = model
        params, chunk_after, self.hook_layers = model.grad_ready_order(n_chunks)
        self.reducer = FlatGradReducer(params, chunk_after=chunk_after)
        act = getattr(model, "act_dtype", torch.float32)
        self.opt = FlatAdam(self.reducer, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay,
                            shadow_dtype=act if (shadow16 and act != torch.float32) else None)
        model._grad_sync = (self.reducer, self.hook_layers)
        self.loss_name = loss
        self.timing: Optional[List] = None     # when a list: (start, end) CUDA events around the exposed part of the all-reduce
        self.cuda_graph = cuda_graph
        self._graph = None                     # (key, CUDAGraph, static x, static labels, static loss)
        # fp16 activations: the reference's recipe relies on accelerate's GradScaler (--mixed_precision=fp16); here a static
        # scale - the loss is multiplied by it before backward (so the 16-bit gradient terms stay in fp16's range), the
        # fused Adam divides it out again, and an eagerly launched step whose gradients overflowed is skipped (p, m, v and
        # the step count untouched; `last_step_applied` tells the caller to lower the scale).  bf16 / fp32: leave it at 1.
        self.loss_scale = float(loss_scale)
        self.last_step_applied = True

    def loss_fn(self, logits, labels):
        if self.loss_name == "bce":
            return torch.nn.functional.binary_cross_entropy_with_logits(logits, labels)
        return torch.nn.functional.cross_entropy(logits, torch.argmax(labels.long(), dim=1))     # (traintest.py:150)

    def _eager(self, x: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        self.reducer.zero()                                   # optimizer.zero_grad()   (traintest.py:167)
        loss = self.loss_fn(self.model(x), labels)            # (:144-152)
        scaled = self.loss_scale != 1.0
        (loss * self.loss_scale if scaled else loss).backward()   # accelerator.backward: chunks launch from hooks (:168)
        if self.timing is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        self.reducer.reduce()                                 # the tail chunk + join
        if self.timing is not None:
            e1.record()
            self.timing.append((e0, e1))
        # (the non-finite check is a host read: not inside a captured graph)
        check = scaled and not torch.cuda.is_current_stream_capturing()
        self.last_step_applied = self.opt.step(grad_scale=1.0 / self.loss_scale, skip_nonfinite=check)      # (:169)
        return loss

    def _capture(self, x: torch.Tensor, labels: torch.Tensor, key):
        from . import mixer
        dev = x.device
        sx, sy = x.clone(), labels.clone()
        opt = self.opt
        # two eager steps on a side stream (allocator pools, kernel attributes, NCCL buffers) - and undone afterwards, so
        # that capturing does not train on the first batch three times
        snap = (opt.flat_p.clone(), opt.m.clone(), opt.v.clone(), opt.step_dev.clone(), opt.t)
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(2):
                self._eager(sx, sy)
        cur.wait_stream(side)
        with torch.no_grad():
            opt.flat_p.copy_(snap[0]); opt.m.copy_(snap[1]); opt.v.copy_(snap[2]); opt.step_dev.copy_(snap[3])
        opt.t = snap[4]
        opt.resync_shadow()
        mixer.bump_generation()              # every derived-weight entry is stale: the capture re-derives (and records) them
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        timing, self.timing = self.timing, None
        try:
            with torch.cuda.graph(g):
                sloss = self._eager(sx, sy)
        finally:
            self.timing = timing
        opt.t = snap[4]                      # the captured opt.step() did its host bookkeeping, but no kernel has run yet
        self._graph = (key, g, sx, sy, sloss)

    def __call__(self, x: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        if not self.cuda_graph or self.timing is not None:
            return self._eager(x, labels)
        key = (tuple(x.shape), x.dtype, tuple(labels.shape), labels.dtype, x.device)
        if self._graph is None or self._graph[0] != key:
            try:
                self._capture(x, labels, key)
            except Exception as e:      # e.g. a collective backend that cannot be captured: keep training, eagerly
                import warnings
                warnings.warn(f"TrainStep: CUDA-graph capture of the training step failed ({type(e).__name__}: {e}); "
                              "falling back to eager launches")
                self.cuda_graph, self._graph = False, None
                torch.cuda.synchronize(x.device)
                return self._eager(x, labels)
        _, g, sx, sy, sloss = self._graph
        sx.copy_(x, non_blocking=True)
        sy.copy_(labels, non_blocking=True)
        g.replay()
        self.opt.after_device_step()         # host mirror of the update the replay just enqueued
        return sloss
