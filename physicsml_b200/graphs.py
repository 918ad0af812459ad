"""Whole-step CUDA-graph capture.

A MACE training step on molecule batches is launch-bound: a few hundred kernels of tens of microseconds each.
Capturing forward + forces + loss + backward (+ gradient all-reduce) + optimiser into one CUDA graph removes the host
from the loop ("CUDA streams and graphs instead of a tracing compiler").  Inputs live in static device buffers; a
new batch of the same shape is copied into them before each replay.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch


class GraphedStep:
    def __init__(self, model: torch.nn.Module, optimizer: torch.optim.Optimizer, data: dict, targets: dict,
                 loss_fn: Callable, after_backward: Optional[Callable] = None, warmup: int = 3) -> None:
        self.model, self.optimizer, self.loss_fn, self.after_backward = model, optimizer, loss_fn, after_backward
        self.data = {k: (v.clone() if torch.is_tensor(v) and v.is_cuda else v) for k, v in data.items()}
        self.targets = {k: v.clone() for k, v in targets.items()}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        self.optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph):
            self.loss = self._body(zero=False)

    def _body(self, zero: bool = True) -> torch.Tensor:
        if zero:
            self.optimizer.zero_grad(set_to_none=True)
        loss = self.loss_fn(self.model, self.data, self.targets)
        loss.backward()
        if self.after_backward is not None:
            self.after_backward()
        self.optimizer.step()
        return loss.detach()

    def load(self, data: dict, targets: dict) -> None:
        """copy a new batch of identical shapes into the static buffers (async on the current stream)"""
        for k, v in data.items():
            if torch.is_tensor(v) and torch.is_tensor(self.data.get(k)) and self.data[k].is_cuda:
                self.data[k].copy_(v, non_blocking=True)
        for k, v in targets.items():
            self.targets[k].copy_(v, non_blocking=True)

    def __call__(self) -> torch.Tensor:
        self.graph.replay()
        return self.loss
