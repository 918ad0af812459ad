"""Fused Adam(amsgrad) over all parameter tensors of all param groups in one kernel.

Drop-in for ``torch.optim.Adam(params, lr, betas, eps, weight_decay, amsgrad=True)`` as the reference configures it
(``models/mace/supervised/default_configs.py:25-33``; one param group PER PARAMETER with its own ``weight_decay``,
``mace_module.py:131-157`` — see :func:`reference_param_groups`): same state names (``exp_avg``, ``exp_avg_sq``,
``max_exp_avg_sq``, ``step``) and one step counter per parameter, so optimizer state dicts interchange with
``torch.optim.Adam`` in both directions; same update rule; CUDA-graph capturable (step counters live on the device).
CUDA only; there is no fallback.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import ops


class _Tensor(C.Structure):
    _fields_ = [("p", C.c_void_p), ("g", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("vmax", C.c_void_p),
                ("n", C.c_longlong), ("step", C.c_void_p), ("lr", C.c_float), ("b1", C.c_float), ("b2", C.c_float),
                ("omb1", C.c_float), ("omb2", C.c_float), ("eps", C.c_float), ("wd", C.c_float), ("gscale", C.c_float)]


class _Chunk(C.Structure):
    _fields_ = [("tensor", C.c_int), ("pad", C.c_int), ("offset", C.c_longlong)]


CHUNK = 4096


def reference_param_groups(module: torch.nn.Module, weight_decay: float, backbone: str = "mace"):
    """The param groups of the reference's ``configure_optimizers`` (models/mace/supervised/mace_module.py:131-157): one
    group per named parameter; ``weight_decay`` only for names collected from ``interactions`` / ``messages`` /
    ``node_updates`` that contain ``linear.weight`` or ``residual_connection_layer.weight``.

    Faithful to the reference INCLUDING its name quirk: the decay list holds names relative to the sub-module
    (``"0.linear.weight"``) while the membership test uses names relative to the whole module
    (``"mace.interactions.0.linear.weight"``), so as shipped no parameter ever matches and every group gets
    ``weight_decay = 0``.  ``match_full_names=True`` semantics are available through :func:`decay_param_groups`."""
    decay = []
    bb = getattr(module, backbone)
    for part in ("interactions", "messages", "node_updates"):
        for name, _ in getattr(bb, part).named_parameters():
            if "linear.weight" in name or "residual_connection_layer.weight" in name:
                decay.append(name)
    return [{"name": name, "params": [p], "weight_decay": weight_decay if name in decay else 0.0}
            for name, p in module.named_parameters() if p.numel() > 0]


def decay_param_groups(module: torch.nn.Module, weight_decay: float, backbone: str = "mace"):
    """what mace_module.py:131-157 evidently intends: decay on the ``linear`` / ``residual_connection_layer`` weights of
    the interaction, message and node-update blocks, none elsewhere (one group per parameter)."""
    out = []
    for name, p in module.named_parameters():
        if p.numel() == 0:
            continue
        inside = any(name.startswith(f"{backbone}.{part}.") for part in ("interactions", "messages", "node_updates"))
        hit = inside and ("linear.weight" in name or "residual_connection_layer.weight" in name)
        out.append({"name": name, "params": [p], "weight_decay": weight_decay if hit else 0.0})
    return out


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 amsgrad: bool = True, grad_scale: float = 1.0) -> None:
        if not amsgrad:
            raise NotImplementedError("FusedAdam implements the amsgrad variant the reference trains with")
        # capturable=True: Optimizer.load_state_dict then moves a loaded `step` to the parameter's device as fp32
        # (torch.optim.Adam keeps it on the CPU otherwise), which is what the kernel dereferences
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=True, capturable=True))
        self.grad_scale = float(grad_scale)  # multiplies every gradient inside the kernel (1/world after a summed all-reduce)
        self.grad_views: dict = {}  # parameter -> tensor to read its gradient from (parallel.GradientReducer.attach)
        self._table = None  # (signature, device bytes, offsets, n chunks, n steps, pinned host copy)

    def _state(self, p):
        st = self.state[p]
        if not st:
            st["step"] = torch.zeros((), dtype=torch.float32, device=p.device)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["max_exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        step = st["step"]
        if not torch.is_tensor(step) or step.device != p.device or step.dtype != torch.float32 or step.dim() != 0:
            # a state dict saved by torch.optim.Adam (CPU / python-number step) or loaded with map_location="cpu"
            st["step"] = torch.as_tensor(float(step), dtype=torch.float32).to(p.device).reshape(())
        for k in ("exp_avg", "exp_avg_sq", "max_exp_avg_sq"):
            if st[k].device != p.device or st[k].dtype != torch.float32 or not st[k].is_contiguous():
                st[k] = st[k].to(device=p.device, dtype=torch.float32).contiguous()
        return st

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        items = []
        for group in self.param_groups:
            for p in group["params"]:
                g = self.grad_views.get(p, p.grad)  # a reduced flat-bucket view takes precedence over .grad
                if g is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous() or not g.is_contiguous():
                    raise TypeError("FusedAdam needs contiguous fp32 CUDA parameters and gradients")
                items.append((p, self._state(p), group, g))
        if not items:
            return loss
        sig = tuple((p.data_ptr(), gr.data_ptr(), st["step"].data_ptr(), st["exp_avg"].data_ptr(), float(g["lr"]),
                     float(g["weight_decay"]), tuple(g["betas"]), float(g["eps"])) for p, st, g, gr in items) + (self.grad_scale,)
        tab = self._table
        if tab is None or tab[0] != sig:
            # (re)build the device tables; under CUDA-graph capture the copy becomes a memcpy node whose pinned source
            # stays alive in self._table
            ts = (_Tensor * len(items))()
            chunks, steps = [], []
            for i, (p, st, g, gr) in enumerate(items):
                b1, b2 = g["betas"]
                ts[i] = _Tensor(p.data_ptr(), gr.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                                st["max_exp_avg_sq"].data_ptr(), p.numel(), st["step"].data_ptr(), float(g["lr"]), float(b1),
                                float(b2), 1.0 - float(b1), 1.0 - float(b2), float(g["eps"]), float(g["weight_decay"]),
                                self.grad_scale)
                chunks += [(i, off) for off in range(0, p.numel(), CHUNK)]
                if st["step"].data_ptr() not in steps:
                    steps.append(st["step"].data_ptr())
            cs = (_Chunk * len(chunks))(*[_Chunk(i, 0, off) for i, off in chunks])
            sp = (C.c_void_p * len(steps))(*steps)
            raw_t, raw_c, raw_s = bytes(ts), bytes(cs), bytes(sp)
            pad = (-len(raw_t)) % 16
            raw = raw_t + b"\0" * pad + raw_c + raw_s
            host = torch.frombuffer(bytearray(raw), dtype=torch.uint8).pin_memory()
            dev = torch.empty(len(raw), dtype=torch.uint8, device=items[0][0].device)
            dev.copy_(host, non_blocking=True)
            tab = (sig, dev, (len(raw_t) + pad, len(raw_t) + pad + len(raw_c)), len(chunks), len(steps), host)
            self._table = tab
        base = tab[1].data_ptr()
        ops._call("pml_adam_amsgrad_groups", base, base + tab[2][0], tab[3], base + tab[2][1], tab[4], ops._stream(), launches=2)
        return loss
