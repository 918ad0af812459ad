"""Whole-step CUDA-graph capture.

A MACE training step (or an energy+forces evaluation) on molecule batches is launch-bound: a few hundred kernels of
tens of microseconds each.  Capturing forward + forces (+ loss + backward + gradient all-reduce + optimiser) into one
CUDA graph removes the host from the loop ("CUDA streams and graphs instead of a tracing compiler").  Inputs live in
static device buffers; a new batch of the same shape is copied into them before each replay.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch


def _static_copy(d: dict) -> dict:
    out = {}
    for k, v in d.items():
        if torch.is_tensor(v) and v.is_cuda:
            out[k] = v.clone()
        elif torch.is_tensor(v) or isinstance(v, (int, float, bool, str)) or v is None:
            out[k] = v
        else:
            # derived Python objects (e.g. a cached EdgePlan) refer to the caller's tensors, not to the static buffers:
            # they would silently go stale on load(); the body rebuilds what it needs from the static tensors
            raise TypeError(f"graph-captured batches may hold tensors and scalars only; got {type(v).__name__} under '{k}'")
    return out


def _load(static: dict, new: dict) -> None:
    for k, v in new.items():
        if torch.is_tensor(v) and torch.is_tensor(static.get(k)) and static[k].is_cuda:
            if static[k].shape != v.shape:
                raise ValueError(f"'{k}': a captured graph replays fixed shapes ({tuple(static[k].shape)} != {tuple(v.shape)})")
            static[k].copy_(v, non_blocking=True)


def _device_of(d: dict):
    for v in d.values():
        if torch.is_tensor(v) and v.is_cuda:
            return v.device
    return torch.device("cuda", torch.cuda.current_device())


class _Captured:
    def _run(self, captured: bool = False):
        """the body inside the zero-arena scope: the ~110 zero-initialised outputs of a step come out of one buffer that
        is cleared by one memset (ops.ZeroArena); only the declared outputs outlive a replay"""
        from . import ops

        with ops.zero_arena(self._arena, self._device):
            return self._body(captured=captured)

    def release(self) -> None:
        """drop the graph and the zero arena.  Gradients of a captured training step live in the arena: they are cleared"""
        self.graph = None
        self.out = None
        model = getattr(self, "model", None)
        if model is not None and getattr(self, "optimizer", None) is not None:
            for p in model.parameters():
                p.grad = None
        arena = getattr(self, "_arena", None)
        if arena is not None and getattr(self, "_owns_arena", True):
            arena.release()

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    def _capture(self, warmup: int, pre_capture: Optional[Callable] = None, arena=None) -> None:
        from . import ops

        self._owns_arena = arena is None
        self._arena = arena if arena is not None else ops.ZeroArena()
        self._device = _device_of(self.data)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 2)):  # at least two eager passes: the first sizes the zero arena, the second uses it
                self._run()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        if pre_capture is not None:
            pre_capture()
        with torch.cuda.graph(self.graph):
            self.out = self._run(captured=True)

    def __call__(self):
        self.graph.replay()
        return self.out


class GraphedStep(_Captured):
    """one training step: zero_grad, loss_fn(model, data, targets), backward, after_backward(), optimizer.step()"""

    def __init__(self, model: torch.nn.Module, optimizer: torch.optim.Optimizer, data: dict, targets: dict,
                 loss_fn: Callable, after_backward: Optional[Callable] = None, warmup: int = 3) -> None:
        self.model, self.optimizer, self.loss_fn, self.after_backward = model, optimizer, loss_fn, after_backward
        self.data, self.targets = _static_copy(data), _static_copy(targets)
        self._capture(warmup, pre_capture=lambda: self.optimizer.zero_grad(set_to_none=True))
        self.loss = self.out

    def _body(self, captured: bool = False) -> torch.Tensor:
        if not captured:
            self.optimizer.zero_grad(set_to_none=True)
        loss = self.loss_fn(self.model, self.data, self.targets)
        loss.backward()
        if self.after_backward is not None:
            self.after_backward()
        self.optimizer.step()
        return loss.detach()

    def load(self, data: dict, targets: dict) -> None:
        """copy a new batch of identical shapes into the static buffers (async on the current stream)"""
        _load(self.data, data)
        _load(self.targets, targets)


class GraphedEval(_Captured):
    """one evaluation ``fn(model, data) -> tensor or tuple of tensors`` (e.g. energy + forces) on static buffers"""

    def __init__(self, model: torch.nn.Module, data: dict, fn: Callable, warmup: int = 3, arena=None) -> None:
        """arena: a caller-owned ops.ZeroArena shared by successive captures (the MD loop re-captures per neighbour list)"""
        self.model, self.fn = model, fn
        self.data = _static_copy(data)
        self._capture(warmup, arena=arena)

    def _body(self, captured: bool = False):
        out = self.fn(self.model, self.data)
        return tuple(o.detach() for o in out) if isinstance(out, (tuple, list)) else out.detach()

    def load(self, data: dict) -> None:
        _load(self.data, data)
