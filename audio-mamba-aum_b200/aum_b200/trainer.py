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
    (traintest.py:32-34: betas (0.95, 0.999), weight_decay 5e-7, lr from the experiment script)."""

    def __init__(self, model, lr: float = 1e-5, betas=(0.95, 0.999), eps: float = 1e-8, weight_decay: float = 5e-7,
                 n_chunks: int = 3, loss: str = "bce"):
        self.model = model
        params, chunk_after, self.hook_layers = model.grad_ready_order(n_chunks)
        self.reducer = FlatGradReducer(params, chunk_after=chunk_after)
        self.opt = FlatAdam(self.reducer, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        model._grad_sync = (self.reducer, self.hook_layers)
        self.loss_name = loss
        self.timing: Optional[List] = None     # when a list: (start, end) CUDA events around the exposed part of the all-reduce

    def loss_fn(self, logits, labels):
        if self.loss_name == "bce":
            return torch.nn.functional.binary_cross_entropy_with_logits(logits, labels)
        return torch.nn.functional.cross_entropy(logits, torch.argmax(labels.long(), dim=1))     # (traintest.py:150)

    def __call__(self, x: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        self.reducer.zero()                                   # optimizer.zero_grad()   (traintest.py:167)
        loss = self.loss_fn(self.model(x), labels)            # (:144-152)
        loss.backward()                                       # accelerator.backward: chunks launch from hooks (:168)
        if self.timing is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        self.reducer.reduce()                                 # the tail chunk + join
        if self.timing is not None:
            e1.record()
            self.timing.append((e0, e1))
        self.opt.step()                                       # (:169)
        return loss
