"""Fused Adam(amsgrad) over all parameter tensors in one kernel.

Drop-in for ``torch.optim.Adam(params, lr, betas, eps, weight_decay, amsgrad=True)`` as the reference configures it
(``models/mace/supervised/default_configs.py:25-33``): same state names (``exp_avg``, ``exp_avg_sq``,
``max_exp_avg_sq``, ``step``) so optimizer state dicts interchange, same update rule, CUDA-graph capturable (the step
counter lives on the device).  CUDA only; there is no fallback.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from . import ops


class _Tensor(C.Structure):
    _fields_ = [("p", C.c_void_p), ("g", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("vmax", C.c_void_p), ("n", C.c_longlong)]


class _Chunk(C.Structure):
    _fields_ = [("tensor", C.c_int), ("pad", C.c_int), ("offset", C.c_longlong)]


CHUNK = 4096


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 amsgrad: bool = True) -> None:
        if not amsgrad:
            raise NotImplementedError("FusedAdam implements the amsgrad variant the reference trains with")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=True))
        self._tables = {}  # group index -> (ptr signature, device tensor table, device chunk table, n chunks, pinned host copy)

    def _state(self, p):
        st = self.state[p]
        if not st:
            st["step"] = torch.zeros((), dtype=torch.float32, device=p.device)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["max_exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            for p in ps:
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous() or not p.grad.is_contiguous():
                    raise TypeError("FusedAdam needs contiguous fp32 CUDA parameters and gradients")
            sts = [self._state(p) for p in ps]
            sig = tuple((p.data_ptr(), p.grad.data_ptr()) for p in ps)
            tab = self._tables.get(gi)
            if tab is None or tab[0] != sig:
                # (re)build the device tables; under CUDA-graph capture the copy becomes a memcpy node whose pinned
                # source stays alive in self._tables
                ts = (_Tensor * len(ps))()
                chunks = []
                for i, (p, st) in enumerate(zip(ps, sts)):
                    ts[i] = _Tensor(p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                                    st["max_exp_avg_sq"].data_ptr(), p.numel())
                    chunks += [(i, off) for off in range(0, p.numel(), CHUNK)]
                cs = (_Chunk * len(chunks))(*[_Chunk(i, 0, off) for i, off in chunks])
                raw = bytes(ts) + bytes(cs)
                host = torch.frombuffer(bytearray(raw), dtype=torch.uint8).pin_memory()
                dev = torch.empty(len(raw), dtype=torch.uint8, device=ps[0].device)
                dev.copy_(host, non_blocking=True)
                tab = (sig, dev, C.sizeof(_Tensor) * len(ps), len(chunks), host)
                self._tables[gi] = tab
            # one step counter per group (all tensors of a group step together); mirrored into every state entry
            step = sts[0]["step"]
            for st in sts[1:]:
                st["step"] = step
            b1, b2 = group["betas"]
            ops._call("pml_adam_amsgrad", tab[1].data_ptr(), tab[1].data_ptr() + tab[2], tab[3], step.data_ptr(),
                      float(group["lr"]), float(b1), float(b2), 1.0 - float(b1), 1.0 - float(b2), float(group["eps"]),
                      float(group["weight_decay"]),
                      ops._stream(), launches=2)
        return loss
