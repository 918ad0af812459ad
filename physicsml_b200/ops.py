"""``torch.library`` custom ops (namespace ``physicsml_b200``) over the C ABI, each with an autograd formula whose
backward is again made of these ops, so that forces (first derivative, ``lightning/module.py:28-48`` of the reference)
and force-matching training (second derivative, ``create_graph=True``) never fall back to generic autograd
gather/scatter.

CUDA only.  There is deliberately no CPU implementation and no PyTorch-eager fallback: calling an op with CPU
tensors raises ``NotImplementedError`` from the dispatcher, and a missing shared library raises
``MissingExtensionError`` at the first launch.
"""
from __future__ import annotations

import contextlib
import weakref
import ctypes as C
import os
from typing import List, Optional, Tuple

import torch
from torch import Tensor

from . import _lib, plans

NS = "physicsml_b200"

# count of kernel launches issued through the C ABI (bench.py reports it as `gpu_launches`)
LAUNCHES = [0]


def _p(t: Optional[Tensor]):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _f32c(t: Tensor) -> Tensor:
    if t.dtype != torch.float32:
        raise TypeError(f"physicsml_b200 ops are fp32 only, got {t.dtype}")
    return t.contiguous()


# PML_POISON_OUTPUTS=1: every output buffer a kernel is expected to overwrite completely is pre-filled with NaN, so a
# kernel that leaves part of its output unwritten fails loudly instead of returning whatever the allocator held
POISON_OUTPUTS = [os.environ.get("PML_POISON_OUTPUTS", "0") == "1"]


def _new_empty(t: Tensor, *shape) -> Tensor:
    if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
        shape = tuple(shape[0])
    return t.new_full(shape, float("nan")) if POISON_OUTPUTS[0] else t.new_empty(shape)


def _empty_like(t: Tensor) -> Tensor:
    return torch.full_like(t, float("nan")) if POISON_OUTPUTS[0] else torch.empty_like(t)


# ---- zero-initialised outputs from one arena ---------------------------------------------------------------------------
# A training step zero-initialises ~110 small outputs (atomically accumulated cotangents, split-K results, weight
# gradients): as separate memsets they are ~110 graph nodes of ~1.5 us + dependency latency each (0.15 ms of a 5.6 ms
# config-2 step).  Inside a ``zero_arena`` scope they are carved out of ONE buffer cleared by ONE memset at scope entry.
# Opt-in and scoped on purpose: a tensor handed out here is re-zeroed at the next entry, so the scope is used by the
# CUDA-graph bodies (graphs.py), where nothing but the declared outputs outlives a replay.  Outside a scope (and on
# overflow) plain ``torch.zeros`` is used.  Process-wide, not thread-local: cotangent formulas run on the autograd
# engine's device thread.
# Every tensor gets its OWN (non-owning) storage object over its slice of the buffer: torch.library checks that a custom
# op's outputs share no storage with its inputs or with each other, and slices of one tensor would all share one.  The
# arena therefore keeps its buffers alive until ``release()``; owners (GraphedStep / GraphedEval) release it when they go.
_FROM_PTR = getattr(torch._C, "_construct_storage_from_data_pointer", None)


class ZeroArena:
    ALIGN = 256

    def __init__(self) -> None:
        self.buf: Optional[Tensor] = None
        self.off = 0
        self.want = 0   # bytes requested during the current / last scope (sizes the buffer of the next one)
        self.served = 0
        self._retired: list = []

    def begin(self, device) -> None:
        if _FROM_PTR is None:  # this torch cannot build aliasing-free views: the arena stays off (plain torch.zeros)
            return
        need = int(self.want * 1.25) + (1 << 20)
        if (self.buf is None or self.buf.numel() < need or self.buf.device != device) and not torch.cuda.is_current_stream_capturing():
            if self.buf is not None:
                self._retired.append(self.buf)  # tensors handed out earlier may still point into it
            self.buf = torch.empty(need, dtype=torch.uint8, device=device)
        self.off, self.want = 0, 0
        if self.buf is not None:
            self.buf.zero_()

    def release(self) -> None:
        """drop the buffers; every tensor handed out must be dead by now"""
        self.buf = None
        self._retired.clear()

    def take(self, ref: Tensor, shape, dtype) -> Optional[Tensor]:
        n = 1
        for d in shape:
            n *= int(d)
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        padded = (nbytes + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.want += padded
        if self.buf is None or ref.device != self.buf.device or self.off + padded > self.buf.numel():
            return None
        storage = _FROM_PTR(self.buf.data_ptr() + self.off, self.buf.device, max(nbytes, 1))
        t = torch.empty(0, dtype=dtype, device=self.buf.device).set_(storage, 0, tuple(int(d) for d in shape))
        self.off += padded
        self.served += 1
        return t


_ARENA: list = [None]


@contextlib.contextmanager
def zero_arena(arena: ZeroArena, device):
    prev = _ARENA[0]
    arena.begin(device)
    _ARENA[0] = arena
    try:
        yield arena
    finally:
        _ARENA[0] = prev


def _zeros(ref: Tensor, *shape, dtype=None) -> Tensor:
    if len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size)):
        shape = tuple(shape[0])
    dtype = dtype or ref.dtype
    a = _ARENA[0]
    if a is not None:
        t = a.take(ref, shape, dtype)
        if t is not None:
            return t
    return torch.zeros(shape, dtype=dtype, device=ref.device)


def _zeros_like(t: Tensor) -> Tensor:
    return _zeros(t, tuple(t.shape))


# optional per-kernel device timing (bench.py): name -> list of (tag, start_event, end_event)
KERNEL_TIMING: dict = {}


def enable_kernel_timing(names) -> None:
    KERNEL_TIMING.clear()
    for n in names:
        KERNEL_TIMING[n] = []


# PML_NVTX=1: every C-ABI launch is wrapped in an NVTX range named after its entry point, so nsys / ncu --nvtx timelines
# show the hot path by operator (the reference has no tracing hooks of its own)
NVTX = [os.environ.get("PML_NVTX", "0") == "1"]


def _call(name: str, *args, launches: int = 1, tag=None) -> None:
    """launch through the C ABI; ``tag`` (default: the first argument) labels the launch in KERNEL_TIMING records"""
    LAUNCHES[0] += launches
    if NVTX[0]:
        torch.cuda.nvtx.range_push(name)
        try:
            _lib.check(getattr(_lib.lib(), name)(*args), name)
        finally:
            torch.cuda.nvtx.range_pop()
        return
    rec = KERNEL_TIMING.get(name[:-3] if name.endswith("_ex") else name)  # pml_tp_fwd_ex is timed as pml_tp_fwd
    if rec is None:
        _lib.check(getattr(_lib.lib(), name)(*args), name)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(getattr(_lib.lib(), name)(*args), name)
    e1.record()
    rec.append((args[0] if tag is None else tag, e0, e1))


# ======================================================================================================================
# edge featurisation
# ======================================================================================================================
def _feat_args(pos, snd, rcv, shift, cell, bessel_w, r_max, prefactor, w_scale, p, lmax):
    return (_p(pos), _p(snd), _p(rcv), _p(shift), _p(cell), _p(bessel_w), float(r_max), float(prefactor),
            float(w_scale), int(p), int(bessel_w.numel()), int(lmax))


@torch.library.custom_op(f"{NS}::edge_featurise", mutates_args=(), device_types="cuda")
def edge_featurise(pos: Tensor, snd: Tensor, rcv: Tensor, shift: Optional[Tensor], cell: Optional[Tensor],
                   bessel_w: Tensor, r_max: float, prefactor: float, w_scale: float, p: int,
                   lmax: int) -> Tuple[Tensor, Tensor]:
    pos = _f32c(pos)
    E = snd.numel()
    Y = _new_empty(pos, E, (lmax + 1) ** 2)
    B = _new_empty(pos, E, bessel_w.numel())
    _call("pml_edge_featurise_fwd", *_feat_args(pos, snd, rcv, shift, cell, bessel_w, r_max, prefactor, w_scale, p, lmax),
          _p(Y), _p(B), None, E, _stream())
    return Y, B


@edge_featurise.register_fake
def _(pos, snd, rcv, shift, cell, bessel_w, r_max, prefactor, w_scale, p, lmax):
    E = snd.numel()
    return _new_empty(pos, E, (lmax + 1) ** 2), _new_empty(pos, E, bessel_w.numel())


@torch.library.custom_op(f"{NS}::edge_featurise_bwd", mutates_args=(), device_types="cuda")
def edge_featurise_bwd(pos: Tensor, snd: Tensor, rcv: Tensor, shift: Optional[Tensor], cell: Optional[Tensor],
                       bessel_w: Tensor, r_max: float, prefactor: float, w_scale: float, p: int, lmax: int,
                       gY: Optional[Tensor], gB: Optional[Tensor]) -> Tensor:
    pos = _f32c(pos)
    gpos = _zeros_like(pos)
    gY = None if gY is None else _f32c(gY)
    gB = None if gB is None else _f32c(gB)
    _call("pml_edge_featurise_bwd", *_feat_args(pos, snd, rcv, shift, cell, bessel_w, r_max, prefactor, w_scale, p, lmax),
          _p(gY), _p(gB), _p(gpos), snd.numel(), _stream())
    return gpos


@edge_featurise_bwd.register_fake
def _(pos, snd, rcv, shift, cell, bessel_w, r_max, prefactor, w_scale, p, lmax, gY, gB):
    return _empty_like(pos)


@torch.library.custom_op(f"{NS}::edge_featurise_jvp", mutates_args=(), device_types="cuda")
def edge_featurise_jvp(pos: Tensor, snd: Tensor, rcv: Tensor, shift: Optional[Tensor], cell: Optional[Tensor],
                       bessel_w: Tensor, r_max: float, prefactor: float, w_scale: float, p: int, lmax: int,
                       t: Tensor) -> Tuple[Tensor, Tensor]:
    pos, t = _f32c(pos), _f32c(t)
    E = snd.numel()
    tY = _new_empty(pos, E, (lmax + 1) ** 2)
    tB = _new_empty(pos, E, bessel_w.numel())
    _call("pml_edge_featurise_jvp", *_feat_args(pos, snd, rcv, shift, cell, bessel_w, r_max, prefactor, w_scale, p, lmax),
          _p(t), _p(tY), _p(tB), E, _stream())
    return tY, tB


@edge_featurise_jvp.register_fake
def _(pos, snd, rcv, shift, cell, bessel_w, r_max, prefactor, w_scale, p, lmax, t):
    E = snd.numel()
    return _new_empty(pos, E, (lmax + 1) ** 2), _new_empty(pos, E, bessel_w.numel())


@torch.library.custom_op(f"{NS}::bessel_weight_grad", mutates_args=(), device_types="cuda")
def bessel_weight_grad(pos: Tensor, snd: Tensor, rcv: Tensor, shift: Optional[Tensor], cell: Optional[Tensor],
                       bessel_w: Tensor, r_max: float, prefactor: float, w_scale: float, p: int, lmax: int,
                       gB: Tensor, t: Optional[Tensor]) -> Tensor:
    """gradient w.r.t. trainable Bessel frequencies (reference nequip/modules/radial.py:9-19): of <gB, B> (t None) or of
    <gB, dB/dd * u.(t[r]-t[s])> (the force pass / its tangent map); see pml_bessel_weight_grad"""
    pos, gB = _f32c(pos), _f32c(gB)
    t = None if t is None else _f32c(t)
    gw = _zeros(bessel_w, tuple(bessel_w.shape), dtype=torch.float32)
    _call("pml_bessel_weight_grad", _p(pos), _p(snd), _p(rcv), _p(shift), _p(cell), _p(bessel_w), float(r_max),
          float(prefactor), float(w_scale), int(p), int(bessel_w.numel()), _p(gB), _p(t), _p(gw), snd.numel(), _stream())
    return gw


@bessel_weight_grad.register_fake
def _(pos, snd, rcv, shift, cell, bessel_w, r_max, prefactor, w_scale, p, lmax, gB, t):
    return torch.empty_like(bessel_w, dtype=torch.float32)


def _feat_save(ctx, inputs, extra=None):
    ctx.save_for_backward(*inputs[:6], extra)  # pos, snd, rcv, shift, cell, bessel_w (None allowed) + one extra tensor
    ctx.consts = tuple(inputs[6:11])


def _feat_args_from(ctx):
    return tuple(ctx.saved_tensors[:6]) + ctx.consts


def _feat_setup(ctx, inputs, output):
    ctx.set_materialize_grads(False)
    _feat_save(ctx, inputs)


def _feat_bwd(ctx, gY, gB):
    g = gw = None
    if ctx.needs_input_grad[0] and (gY is not None or gB is not None):
        g = edge_featurise_bwd(*_feat_args_from(ctx), gY, gB)
    if ctx.needs_input_grad[5] and gB is not None and _PARAM_GRADS[0]:  # trainable Bessel frequencies
        gw = bessel_weight_grad(*_feat_args_from(ctx), gB, None)
    return (g, None, None, None, None, gw) + (None,) * 5


edge_featurise.register_autograd(_feat_bwd, setup_context=_feat_setup)


def _feat_bwd_setup(ctx, inputs, output):
    ctx.set_materialize_grads(False)
    _feat_save(ctx, inputs, inputs[12] if (inputs[12] is not None and inputs[5].requires_grad) else None)
    ctx.has = (inputs[11] is not None, inputs[12] is not None)


def _feat_bwd_bwd(ctx, v):
    # gpos is linear in (gY, gB): their cotangents are the tangent map applied to v.  The second derivative of the
    # harmonics / radial basis w.r.t. the coordinates only feeds d(loss)/d(coordinates), which nothing consumes
    # (SURVEY §7 "hard parts"); it is not propagated.
    # (Consequence: d(force loss)/d(coordinates) is NOT available from this path; parameter gradients are exact.)
    if v is None:
        return (None,) * 13
    tY, tB = edge_featurise_jvp(*_feat_args_from(ctx), v)
    gw = None
    gB = ctx.saved_tensors[6]
    if ctx.needs_input_grad[5] and gB is not None:  # d<v, gpos>/d bessel_w: gpos is linear in dB/dd
        gw = bessel_weight_grad(*_feat_args_from(ctx), gB, v)
    return (None,) * 5 + (gw,) + (None,) * 5 + (tY if ctx.has[0] else None, tB if ctx.has[1] else None)


edge_featurise_bwd.register_autograd(_feat_bwd_bwd, setup_context=_feat_bwd_setup)


def _feat_jvp_setup(ctx, inputs, output):
    ctx.set_materialize_grads(False)
    _feat_save(ctx, inputs, inputs[11] if inputs[5].requires_grad else None)


def _feat_jvp_bwd(ctx, gtY, gtB):
    g = gw = None
    if gtY is not None or gtB is not None:
        g = edge_featurise_bwd(*_feat_args_from(ctx), gtY, gtB)
    t = ctx.saved_tensors[6]
    if ctx.needs_input_grad[5] and gtB is not None and t is not None:  # tB = dB/dd * u.(t[r]-t[s])
        gw = bessel_weight_grad(*_feat_args_from(ctx), gtB, t)
    return (None,) * 5 + (gw,) + (None,) * 5 + (g,)


edge_featurise_jvp.register_autograd(_feat_jvp_bwd, setup_context=_feat_jvp_setup)


# ======================================================================================================================
# fused tensor product + scatter
# ======================================================================================================================
@torch.library.custom_op(f"{NS}::tp_scatter", mutates_args=(), device_types="cuda")
def tp_scatter(h: Tensor, Y: Tensor, R: Tensor, gather: Tensor, rowptr: Tensor, perm: Optional[Tensor],
               plan: int) -> Tensor:
    pl = plans.get(plan)
    h, Y, R = _f32c(h), _f32c(Y), _f32c(R)
    N = rowptr.numel() - 1
    out = _new_empty(h, N, pl.C * pl.DOUT)
    _call("pml_tp_fwd_ex", pl.spec_id, _p(h), _p(Y), _p(R), _p(gather), _p(rowptr), _p(perm), _p(out), N, pl.C,
          pl.out_layout_id, _stream())
    return out


@tp_scatter.register_fake
def _(h, Y, R, gather, rowptr, perm, plan):
    pl = plans.get(plan)
    return _new_empty(h, rowptr.numel() - 1, pl.C * pl.DOUT)


@torch.library.custom_op(f"{NS}::tp_scatter_bwd", mutates_args=(), device_types="cuda")
def tp_scatter_bwd(h: Tensor, Y: Tensor, R: Tensor, g: Tensor, gather: Tensor, rowptr: Tensor, perm: Optional[Tensor],
                   plan: int, need_h: bool, need_y: bool, need_r: bool, anchor: Optional[Tensor] = None
                   ) -> Tuple[Tensor, Tensor, Tensor]:
    """anchor: not read.  The first-order formula passes the forward's OUTPUT here when it records this op (create_graph), which
    makes the forward's node a graph successor of this op's node: a backward pass that reaches this node then also runs
    the forward's cotangent formula, AFTER this one - what the in-kernel accumulation of the two dR contributions relies
    on (see _tp_bwd_bwd)."""
    pl = plans.get(plan)
    h, Y, R, g = _f32c(h), _f32c(Y), _f32c(R), _f32c(g)
    N = rowptr.numel() - 1
    dh = _zeros_like(h) if need_h else _new_empty(h, 0)
    dY = (_zeros_like(Y) if (pl.C > 32 or pl.n_groups > 1) else _empty_like(Y)) if need_y else _new_empty(Y, 0)
    dR = _empty_like(R) if need_r else _new_empty(R, 0)
    if need_h or need_y or need_r:
        _call("pml_tp_bwd_ex", pl.spec_id, _p(h), _p(Y), _p(R), _p(g), _p(gather), _p(rowptr), _p(perm),
              _p(dh) if need_h else None, _p(dY) if need_y else None, _p(dR) if need_r else None, N, pl.C,
              pl.out_layout_id, _stream(), tag=(pl.spec_id, need_h, need_y, need_r))
    return dh, dY, dR


@tp_scatter_bwd.register_fake
def _(h, Y, R, g, gather, rowptr, perm, plan, need_h, need_y, need_r, anchor=None):
    return (_empty_like(h) if need_h else _new_empty(h, 0), _empty_like(Y) if need_y else _new_empty(Y, 0),
            _empty_like(R) if need_r else _new_empty(R, 0))


@torch.library.custom_op(f"{NS}::tp_scatter_bwd_acc", mutates_args=("dR",), device_types="cuda")
def tp_scatter_bwd_acc(h: Tensor, Y: Tensor, R: Tensor, g: Tensor, gather: Tensor, rowptr: Tensor, perm: Optional[Tensor],
                       plan: int, need_h: bool, need_y: bool, dR: Tensor) -> Tuple[Tensor, Tensor]:
    """tp_scatter_bwd whose R cotangent is ADDED into the existing ``dR`` (in place, by L2 reductions inside the kernel)
    instead of returned: the second-order pass produces two dR contributions, and adding them with a separate kernel
    moved 3 x E x P x C x 4 bytes (640 MB at config 2, 3.6 GB at config 5).  Returns (dh, dY).  No autograd formula (it
    would be third order)."""
    pl = plans.get(plan)
    h, Y, R, g = _f32c(h), _f32c(Y), _f32c(R), _f32c(g)
    if not dR.is_contiguous() or dR.dtype != torch.float32:
        raise TypeError("dR must be a contiguous fp32 tensor")
    N = rowptr.numel() - 1
    dh = _zeros_like(h) if need_h else _new_empty(h, 0)
    dY = (_zeros_like(Y) if (pl.C > 32 or pl.n_groups > 1) else _empty_like(Y)) if need_y else _new_empty(Y, 0)
    _call("pml_tp_bwd_ex", pl.spec_id, _p(h), _p(Y), _p(R), _p(g), _p(gather), _p(rowptr), _p(perm),
          _p(dh) if need_h else None, _p(dY) if need_y else None, _p(dR), N, pl.C, pl.out_layout_id | 2, _stream(),
          tag=(pl.spec_id, need_h, need_y, True))
    return dh, dY


@tp_scatter_bwd_acc.register_fake
def _(h, Y, R, g, gather, rowptr, perm, plan, need_h, need_y, dR):
    return (_empty_like(h) if need_h else _new_empty(h, 0), _empty_like(Y) if need_y else _new_empty(Y, 0))


def _tp_setup(ctx, inputs, output):
    h, Y, R, gather, rowptr, perm, plan = inputs
    ctx.save_for_backward(h, Y, R, gather, rowptr, perm, output)
    ctx.plan = plan


# PML_FUSE_DR=0: leave the sum of the two R cotangents of a force-matching step to the autograd engine (A/B runs)
FUSE_DR = [os.environ.get("PML_FUSE_DR", "1") == "1"]
# R cotangent of a second-order tensor-product node, waiting for the first-order node of the same R (keyed by R's storage)
_PENDING_DR: dict = {}


def reset_pending() -> None:
    """drop anything a backward pass that died half-way (an exception inside a formula) left parked; called by the models at
    the start of every forward, which never runs inside a backward pass"""
    if _PENDING_DR:
        _PENDING_DR.clear()


def _pending_dr_check() -> None:
    """end of a backward pass.  A contribution that is still parked belongs to a tensor product whose forward node the engine
    did not execute - which, given the graph edge through `anchor`, only happens when the pass was restricted
    (``autograd.grad(..., inputs=...)``) to inputs that none of h, Y, R lead to: nobody asked for that gradient.
    PML_FUSE_DR_STRICT=1 turns the silent drop into an error (debugging aid)."""
    if _PENDING_DR:
        n = len(_PENDING_DR)
        _PENDING_DR.clear()
        if os.environ.get("PML_FUSE_DR_STRICT") == "1":
            raise RuntimeError(f"physicsml_b200: {n} parked tensor-product weight cotangent(s) were never consumed")


def _tp_bwd(ctx, g):
    h, Y, R, gather, rowptr, perm, out = ctx.saved_tensors
    nh, ny, nr = ctx.needs_input_grad[:3]
    if nr and _PENDING_DR and not torch.is_grad_enabled():
        parked = _PENDING_DR.pop((R.data_ptr(), tuple(R.shape)), None)
        if parked is not None:
            # the second-order node of this tensor product already produced its share of dR: add this one onto it inside
            # the kernel (red.global.add) instead of handing the engine two R-sized tensors to sum
            dh, dY = tp_scatter_bwd_acc(h, Y, R, g, gather, rowptr, perm, ctx.plan, nh, ny, parked)
            return (dh if nh else None, dY if ny else None, parked, None, None, None, None)
    anchor = out if (torch.is_grad_enabled() and nr and FUSE_DR[0]) else None
    dh, dY, dR = tp_scatter_bwd(h, Y, R, g, gather, rowptr, perm, ctx.plan, nh, ny, nr, anchor)
    return (dh if nh else None, dY if ny else None, dR if nr else None, None, None, None, None)


tp_scatter.register_autograd(_tp_bwd, setup_context=_tp_setup)


def _tp_bwd_setup(ctx, inputs, output):
    h, Y, R, g, gather, rowptr, perm, plan, need_h, need_y, need_r, anchor = inputs
    ctx.set_materialize_grads(False)
    ctx.save_for_backward(h, Y, R, g, gather, rowptr, perm)
    ctx.plan = plan
    ctx.needs = (need_h, need_y, need_r)
    ctx.anchor_shape = None if anchor is None else tuple(anchor.shape)


def _acc(a, b):
    if a is None:
        return b
    if b is None:
        return a
    return a + b


def _tp_bwd_bwd(ctx, vh, vY, vR):
    """The op is the gradient of the 4-linear form Phi(h, Y, R, g); differentiating
    <vh, dPhi/dh> + <vY, dPhi/dY> + <vR, dPhi/dR> needs only the same two kernels with permuted arguments."""
    h, Y, R, g, gather, rowptr, perm = ctx.saved_tensors
    plan = ctx.plan
    nh, ny, nr, ng = ctx.needs_input_grad[:4]
    if not ctx.needs[0]:
        vh = None
    if not ctx.needs[1]:
        vY = None
    if not ctx.needs[2]:
        vR = None
    gh = gY = gR = gg = None
    if vh is not None:  # Phi(vh, Y, R, g)
        if ng:
            gg = _acc(gg, tp_scatter(vh, Y, R, gather, rowptr, perm, plan))
        if ny or nr:
            _, a, b = tp_scatter_bwd(vh, Y, R, g, gather, rowptr, perm, plan, False, ny, nr)
            gY, gR = _acc(gY, a if ny else None), _acc(gR, b if nr else None)
    if vY is not None:  # Phi(h, vY, R, g)
        if ng:
            gg = _acc(gg, tp_scatter(h, vY, R, gather, rowptr, perm, plan))
        if nr and gR is not None and not torch.is_grad_enabled():
            # second contribution to dR: accumulated inside the kernel onto the first one (no R-sized add kernel)
            a, _ = tp_scatter_bwd_acc(h, vY, R, g, gather, rowptr, perm, plan, nh, False, gR)
            gh = _acc(gh, a if nh else None)
        elif nh or nr:
            a, _, b = tp_scatter_bwd(h, vY, R, g, gather, rowptr, perm, plan, nh, False, nr)
            gh, gR = _acc(gh, a if nh else None), _acc(gR, b if nr else None)
    if vR is not None:  # Phi(h, Y, vR, g)
        if ng:
            gg = _acc(gg, tp_scatter(h, Y, vR, gather, rowptr, perm, plan))
        if nh or ny:
            a, b, _ = tp_scatter_bwd(h, Y, vR, g, gather, rowptr, perm, plan, nh, ny, False)
            gh, gY = _acc(gh, a if nh else None), _acc(gY, b if ny else None)
    if (gR is not None and ctx.anchor_shape is not None and FUSE_DR[0] and not torch.is_grad_enabled()
            and ctx.needs_input_grad[11]):
        # Force-matching training: R gets one cotangent from this node and one from the forward's node.  Park this one; the
        # forward's formula adds its own onto it inside the kernel and returns the sum.  It runs later in the same pass
        # because it is a successor of this node through `anchor`: the engine counts that edge as a dependency whatever
        # this formula returns for it (nothing), and calls the forward's formula even if no other cotangent reaches it
        # (with a zero one).  Saves an R-sized aten::add per layer (config 2: 213 + 85 MB tensors; NequIP 0.42 ms of a 3.6 ms
        # training step).
        key = (R.data_ptr(), tuple(R.shape))
        if key not in _PENDING_DR:
            if not _PENDING_DR:
                torch.autograd.Variable._execution_engine.queue_callback(_pending_dr_check)
            _PENDING_DR[key] = gR
            gR = None
    return (gh, gY, gR, gg) + (None,) * 8


tp_scatter_bwd.register_autograd(_tp_bwd_bwd, setup_context=_tp_bwd_setup)


# ======================================================================================================================
# per-irrep linear (grouped GEMM)
# ======================================================================================================================
# "tc": tcgen05 3xTF32 kernel for every problem with >= TC_MIN_ROWS[0] runtime rows (the edge-sized radial-MLP GEMMs,
# their weight gradients and the node-sized per-irrep mixing of a batch; reductions longer than 256 promote the TMEM
# accumulators into fp32 registers so the tensor core's truncating accumulation stays below 1e-6 relative), FFMA
# kernel below that (a handful of rows cannot fill a 128-row MMA tile); "ffma": FFMA only
GEMM_MODE = [os.environ.get("PML_GEMM", "tc")]
TC_MIN_ROWS = [int(os.environ.get("PML_TC_MIN_ROWS", "256"))]


def _gemm_name(n_runtime: int) -> str:
    return "pml_gemm_grouped_tc" if (GEMM_MODE[0] == "tc" and n_runtime >= TC_MIN_ROWS[0]) else "pml_gemm_grouped"


def _gemm(name: str, a, b, c, table, nprob: int, m_runtime: int, max_m: int, max_n: int, max_batch: int, splitk: int,
          max_k: int) -> None:
    if name.endswith("_tc"):
        _call(name, a, b, c, table, nprob, m_runtime, max_m, max_n, max_batch, splitk, max_k, _stream())
    else:
        _call(name, a, b, c, table, nprob, m_runtime, max_m, max_n, max_batch, splitk, _stream())


_SM_COUNT: dict = {}


def _sms() -> int:
    """SM count of the current device (148 on B200; also the value assumed when no device is visible: host-side tests)"""
    if not torch.cuda.is_available():
        return 148
    dev = torch.cuda.current_device()
    if dev not in _SM_COUNT:
        _SM_COUNT[dev] = torch.cuda.get_device_properties(dev).multi_processor_count
    return _SM_COUNT[dev]


def _tc_slots(bn: int, promote: bool) -> int:
    """CTAs of gemm_grouped_tc_kernel<bn, promote> resident on the GPU (registers bound it: 96-126 registers x 256
    threads -> 2 per SM; the 80-register non-promoting BN <= 64 variants -> 3 per SM)"""
    return _sms() * (2 if (promote or bn >= 128) else 3)


def _splitk_for(k_extent: int, ctas_without_split: int, slots: Optional[int] = None) -> int:
    """split the reduction so that the grid fills `slots` CTAs without spilling into a nearly empty extra wave
    (floor, not ceil: 10 tiles x 60 splits = 600 CTAs on 592 slots ran three rounds instead of two), keeping >= 4
    k-tiles (of 16) per CTA"""
    if slots is None:
        slots = 4 * _sms()
    want = slots // max(1, ctas_without_split)
    return max(1, min(want, max(1, k_extent // 64), 4096))


def _splitk_rows(tiles: int, k_extent: int, slots: int) -> int:
    """split-K for a row-tiled GEMM with a long reduction (dx of the wide radial layer: 325 row tiles on 296 slots
    = two rounds, the second one 10 % full).  Cost model: rounds x (k-blocks x 1.1 us + epilogue)."""
    best, best_cost = 1, None
    for sk in range(1, 9):
        if sk > 1 and k_extent // sk < 128:
            break
        rounds = (tiles * sk + slots - 1) // slots
        cost = rounds * (k_extent / (16.0 * sk) * 1.1 + (5.0 if sk > 1 else 3.0))
        if best_cost is None or cost < 0.95 * best_cost:
            best, best_cost = sk, cost
    return best


# Node-sized problems (a batch of molecules: ~10 row tiles) occupy a handful of SMs and are latency chains: ~1 us per
# 16-wide k-block whatever the tile multiplies (scripts/bench_small_gemm.py).  Long reductions (K >= 256) are split over
# otherwise idle SMs, with an atomic epilogue into a zeroed output: 1 304 x 256 x 128: 17.1 -> 11.4 us, x 512: 30.4 ->
# 12.9 us.  Below that the zero-fill and the atomic epilogue cost what the shorter chain saves (K = 64: 8.5 -> 11.3 us).
# PML_SMALL_SPLITK=0 switches it off (A/B runs).
SMALL_SPLITK = [os.environ.get("PML_SMALL_SPLITK", "1") == "1"]


def _splitk_small(tiles: int, k_extent: int) -> int:
    if not SMALL_SPLITK[0] or k_extent < 256 or tiles * 4 > _sms():
        return 1
    return max(1, min(k_extent // 64, _sms() // tiles, 8))


# PML_COMPONENT_COPIES=0: let the tensor-core GEMM read e3nn-ordered operands in place (strided reduction index; A/B runs)
COMPONENT_COPIES = [os.environ.get("PML_COMPONENT_COPIES", "1") == "1"]

# PML_TRANSPOSED_DW=0: compute the weight gradient of few-in / many-out layers as dW instead of dW^T (A/B runs)
TRANSPOSED_DW = [os.environ.get("PML_TRANSPOSED_DW", "1") == "1"]

WIDE_MIN_ROWS = [int(os.environ.get("PML_WIDE_MIN_ROWS", "4096"))]


def _wide_ok(pl, n: int) -> bool:
    """the streaming kernels of csrc/gemm_wide.cu take the wide scalar layers (radial-MLP output) on edge-sized inputs"""
    return pl.wide is not None and GEMM_MODE[0] == "tc" and n >= WIDE_MIN_ROWS[0]


NARROW_MIN_ROWS = [int(os.environ.get("PML_NARROW_MIN_ROWS", "262144"))]


def _narrow_ok(pl, n: int) -> bool:
    """64 -> 64 scalar layers on very large edge sets (a periodic box: 2 M edges) also go through the streaming dx
    kernel; on a batch of molecules the extra pack launch costs more than the kernel saves"""
    return pl.narrow is not None and GEMM_MODE[0] == "tc" and n >= NARROW_MIN_ROWS[0]


def _stream_reduce(a: Tensor, w: Tensor, rows: int, n_red: int, n_out: int, w_rs: int, w_cs: int, alpha: float) -> Tensor:
    """out[rows, n_out] = alpha * a[rows, n_red] @ Wt^T with Wt[o, r] = w[o*w_rs + r*w_cs] (streaming dx kernel)"""
    out = _new_empty(a, rows, n_out)
    image = torch.empty(((n_red + 63) // 64) * 2 * 64 * 64, device=w.device, dtype=torch.float32)
    _call("pml_wide_pack_wt", _p(w), w_rs, w_cs, n_out, n_red, _p(image), _stream())
    _call("pml_wide_bwd_x", _p(a), n_red, n_red, _p(image), _p(out), n_out, n_out, rows, alpha, _stream())
    return out


def _wide_image(w: Tensor, k: int, n: int, w_rs: int, w_cs: int) -> Tensor:
    """tf32 hi/lo shared-memory images of W [k, n] (element (i, j) at w[i*w_rs + j*w_cs]), one per 64-column tile"""
    image = torch.empty(((n + 63) // 64) * 2 * 64 * 64, device=w.device, dtype=torch.float32)
    _call("pml_wide_pack_w", _p(w), w_rs, w_cs, k, n, _p(image), _stream())
    return image


@torch.library.custom_op(f"{NS}::irrep_linear", mutates_args=(), device_types="cuda")
def irrep_linear(x: Tensor, w: Tensor, plan: int) -> Tensor:
    pl = plans.get(plan)
    x, w = _f32c(x), _f32c(w)
    n = x.shape[0]
    shape = (n, pl.channels, pl.S) if pl.out_layout == "channel" else (n, pl.d_out)
    if _narrow_ok(pl, n):
        k, nn, alpha = pl.narrow  # y = x W: reduce over k, W[r = k index, o = nn index] at w[r*nn + o]
        return _stream_reduce(x, w, n, k, nn, 1, nn, alpha)
    if _wide_ok(pl, n):
        k, nn, alpha = pl.wide
        out = _new_empty(x, shape)
        image = _wide_image(w, k, nn, nn, 1)
        _call("pml_wide_fwd", _p(x), k, k, _p(image), _p(out), nn, n, nn, alpha, _stream())
        return out
    sk = 1
    if _gemm_name(n).endswith("_tc") and not pl.fwd_serial:
        bn = 32 if pl.max_mul_out <= 32 else (64 if pl.max_mul_out <= 64 else 128)
        sk = _splitk_small(((n + 127) // 128) * ((pl.max_mul_out + bn - 1) // bn) * len(pl._fwd) * pl.max_d, pl.max_k_fwd)
    max_k = pl.max_k_fwd if sk == 1 else (pl.max_k_fwd + sk - 1) // sk + 32
    out = _zeros(x, shape) if (pl.has_unfed_output or sk > 1) else _new_empty(x, shape)
    fwd, _, _ = pl.tables(x.device)
    if fwd is not None and n > 0 and pl._fwd_c and _gemm_name(n).endswith("_tc") and COMPONENT_COPIES[0]:
        # e3nn-ordered input with l > 0 blocks: the GEMM reads a component-major copy (unit-stride reduction index)
        cmap, fwd_c = pl.tables_c(x.device)[0]
        xc = _new_empty(x, n, pl.d_in)
        _call("pml_permute_columns", _p(x), pl.d_in, _p(cmap), _p(xc), pl.d_in, n, _stream())
        _gemm(_gemm_name(n), _p(xc), _p(w), _p(out), _p(fwd_c), len(pl._fwd_c), n, n, pl.max_mul_out, pl.max_d, sk, max_k)
        return out
    if fwd is not None and n > 0 and not pl.fwd_serial:
        _gemm(_gemm_name(n), _p(x), _p(w), _p(out), _p(fwd), len(pl._fwd), n, n, pl.max_mul_out, pl.max_d, sk, max_k)
        return out
    if fwd is not None and n > 0:
        nprob = len(pl._fwd)
        if pl.fwd_serial:
            sz = C.sizeof(_lib.GemmProblem)
            for k in range(nprob):
                _gemm(_gemm_name(n), _p(x), _p(w), _p(out), fwd.data_ptr() + k * sz, 1, n, n, pl.max_mul_out, pl.max_d, 1,
                      pl.max_k_fwd)
        else:
            _gemm(_gemm_name(n), _p(x), _p(w), _p(out), _p(fwd), nprob, n, n, pl.max_mul_out, pl.max_d, 1, pl.max_k_fwd)
    return out


@irrep_linear.register_fake
def _(x, w, plan):
    pl = plans.get(plan)
    if pl.out_layout == "channel":
        return _new_empty(x, x.shape[0], pl.channels, pl.S)
    return _new_empty(x, x.shape[0], pl.d_out)


@torch.library.custom_op(f"{NS}::irrep_linear_bwd_x", mutates_args=(), device_types="cuda")
def irrep_linear_bwd_x(g: Tensor, w: Tensor, plan: int) -> Tensor:
    pl = plans.get(plan)
    g, w = _f32c(g), _f32c(w)
    n = g.shape[0]
    if _narrow_ok(pl, n):
        k, nn, alpha = pl.narrow  # dx = g W^T: reduce over nn, W[o = k index, r = nn index] at w[o*nn + r]
        return _stream_reduce(g, w, n, nn, k, nn, 1, alpha)
    if _wide_ok(pl, n):
        k, nn, alpha = pl.wide  # same reduction as the narrow case, over the wide dimension
        return _stream_reduce(g, w, n, nn, k, nn, 1, alpha)
    name = _gemm_name(n)
    sk = 1
    if name.endswith("_tc") and pl.max_k_bwd_x >= 256:
        bn = 32 if pl.max_mul_in <= 32 else 64
        tiles = ((n + 127) // 128) * ((pl.max_mul_in + bn - 1) // bn) * len(pl._bwd_x) * pl.max_d
        sk = _splitk_rows(tiles, pl.max_k_bwd_x, _tc_slots(bn, True))
    if name.endswith("_tc") and sk == 1:
        bn = 32 if pl.max_mul_in <= 32 else (64 if pl.max_mul_in <= 64 else 128)
        sk = _splitk_small(((n + 127) // 128) * ((pl.max_mul_in + bn - 1) // bn) * len(pl._bwd_x) * pl.max_d, pl.max_k_bwd_x)
    dx = _zeros(g, n, pl.d_in) if (pl.has_unfed_input or sk > 1) else _new_empty(g, n, pl.d_in)
    _, bx, _ = pl.tables(g.device)
    if bx is not None and n > 0 and pl._bwd_x_c and name.endswith("_tc") and COMPONENT_COPIES[0]:
        # cotangent stored in the e3nn order / channel layout: reduce over a component-major copy of it
        cmap, bx_c = pl.tables_c(g.device)[1]
        gc = _new_empty(g, n, pl.d_out)
        _call("pml_permute_columns", _p(g), pl.d_out, _p(cmap), _p(gc), pl.d_out, n, _stream())
        _gemm(name, _p(gc), _p(w), _p(dx), _p(bx_c), len(pl._bwd_x_c), n, n, pl.max_mul_in, pl.max_d, sk,
              (pl.max_k_bwd_x + sk - 1) // sk + 16 if sk > 1 else pl.max_k_bwd_x)
        return dx
    if bx is not None and n > 0:
        _gemm(name, _p(g), _p(w), _p(dx), _p(bx), len(pl._bwd_x), n, n, pl.max_mul_in, pl.max_d, sk,
              (pl.max_k_bwd_x + sk - 1) // sk + 16 if sk > 1 else pl.max_k_bwd_x)
    return dx


@irrep_linear_bwd_x.register_fake
def _(g, w, plan):
    return _new_empty(g, g.shape[0], plans.get(plan).d_in)


@torch.library.custom_op(f"{NS}::irrep_linear_bwd_w", mutates_args=(), device_types="cuda")
def irrep_linear_bwd_w(x: Tensor, g: Tensor, plan: int) -> Tensor:
    pl = plans.get(plan)
    x, g = _f32c(x), _f32c(g)
    n = x.shape[0]
    dw = _zeros(x, pl.weight_numel)
    _, _, bw = pl.tables(x.device)
    if bw is not None and n > 0:
        name = _gemm_name(n)
        A, B, max_m, max_n = x, g, pl.max_mul_in, pl.max_mul_out
        if name.endswith("_tc") and pl._bwd_w_t and TRANSPOSED_DW[0]:
            # few inputs -> many outputs (radial output layer): dW^T = g^T x, the wide dimension on the 128-row tile
            bw, A, B, max_m, max_n = pl.table_bwd_w_t(x.device), g, x, pl.max_mul_out, pl.max_mul_in
        if name.endswith("_tc"):
            bn = 32 if max_n <= 32 else (64 if max_n <= 64 else 128)
            ctas = ((max_m + 127) // 128) * ((max_n + bn - 1) // bn) * len(pl._bwd_w) * pl.max_d
        else:
            ctas = ((max_m + 63) // 64) * ((max_n + 63) // 64) * len(pl._bwd_w) * pl.max_d
        sk = _splitk_for(n, ctas, 2 * _tc_slots(bn, False) if name.endswith("_tc") and bn >= 128 else
                         (_tc_slots(bn, False) if name.endswith("_tc") else 4 * _sms()))
        k_cta = (n + sk - 1) // sk + 16
        if name.endswith("_tc") and k_cta > 1024:
            # Long reductions would run the accumulator-promoting kernel (BN <= 64, 126 registers: two CTAs per SM).  Two
            # candidates, costed in k-blocks (1.3 us promoting, 1.1 us plain) x rounds of resident CTAs, counting only the
            # (problem, batch) pairs that exist: (a) promoting, split sized for ITS tiles and slots, rounded down to
            # whole waves; (b) enough splits to keep every CTA's reduction <= 1008 terms on the plain kernel (config 5,
            # the 17-block weight gradient of `linear` over 1728 nodes: 331 us promoting, one split -> two plain splits).
            tm, nb = (max_m + 127) // 128, pl.bwd_w_batches
            bn_p = 32 if max_n <= 32 else 64
            ctas_p = tm * ((max_n + bn_p - 1) // bn_p) * nb
            sk_p = max(1, (2 * _sms()) // max(1, ctas_p))
            atom = lambda ctas_, bn_: ctas_ * 128.0 * bn_ / 2.0e5  # split-K epilogue: ~2e5 fp32 atomics per microsecond
            cost_p = -(-ctas_p * sk_p // (2 * _sms())) * ((n / sk_p) / 16.0 * 1.3 + 8.0) + atom(ctas_p * sk_p, bn_p)
            sk_n = -(-n // 1008)
            ctas_n = tm * ((max_n + bn - 1) // bn) * nb
            slots_n = _tc_slots(bn, False)
            cost_n = -(-ctas_n * sk_n // slots_n) * ((n / sk_n) / 16.0 * 1.1 + 8.0) + atom(ctas_n * sk_n, bn)
            sk = sk_n if cost_n < cost_p else sk_p
            k_cta = (n + sk - 1) // sk + 16
        # Weight gradients are sums over nodes / edges whose fp32 evaluation order already carries ~sqrt(n) eps; up to
        # 1024 terms per CTA they run on the plain TMEM accumulator (worst-case truncation bias 2e-5 relative, measured
        # < 1e-6 on random data) instead of the promoting variant, which is 1.5x slower on these shapes.
        _gemm(name, _p(A), _p(B), _p(dw), _p(bw), len(pl._bwd_w), n, max_m, max_n, pl.max_d, sk,
              min(k_cta, 256) if k_cta <= 1024 else k_cta)
    return dw


@irrep_linear_bwd_w.register_fake
def _(x, g, plan):
    return _new_empty(x, plans.get(plan).weight_numel)


def _lin_setup(ctx, inputs, output):
    x, w, plan = inputs
    ctx.save_for_backward(x, w)
    ctx.plan = plan


# ``ctx.needs_input_grad`` is fixed when the forward runs (weights require grad), so the cotangent formulas below would
# compute every weight gradient even when the caller differentiates w.r.t. the coordinates only - the force pass of
# every evaluation and of every training step.  PyTorch prunes those branches for its own ops (one node per gradient);
# a fused op cannot be pruned from outside, so ``compute_forces_by_gradient`` says so explicitly.  (Config 4: 19.5 of
# 62.8 ms of an energy+forces evaluation were weight gradients nobody asked for.)
class _ParamGrads:
    """process-wide switch.  NOT thread-local on purpose: autograd evaluates the formulas of a CUDA graph on the engine's
    device worker thread, not on the thread that called ``autograd.grad`` — a thread-local set by the caller would never
    be seen by the formulas (measured: the force pass then computes every weight gradient again, +0.6 ms per config-2
    step).  The price: a backward() started concurrently from ANOTHER Python thread while this context is active would
    also skip its weight gradients; the context is therefore entered only around the force-pass ``autograd.grad`` call."""

    def __init__(self) -> None:
        self.on = True

    def __getitem__(self, _i):
        return self.on


_PARAM_GRADS = _ParamGrads()


@contextlib.contextmanager
def input_grads_only():
    """inside: first-order cotangent formulas skip the parameter gradients (the caller asks for input gradients only,
    e.g. forces = -dE/dx); the recorded graph (create_graph=True) still reaches the parameters through the input
    gradients, which is all force-matching training differentiates.

    Only wrap ``torch.autograd.grad(..., inputs=[coordinates])`` calls in it: a ``backward()`` or a gradient w.r.t.
    parameters taken INSIDE the context would silently miss the weight gradients of the fused ops, which is why the
    context is entered by ``compute_forces_by_gradient`` / ``workloads.training_loss`` themselves and nowhere else."""
    old = _PARAM_GRADS.on
    _PARAM_GRADS.on = False
    try:
        yield
    finally:
        _PARAM_GRADS.on = old


def _lin_bwd(ctx, g):
    x, w = ctx.saved_tensors
    nx, nw = ctx.needs_input_grad[:2]
    nw = nw and _PARAM_GRADS[0]
    return (irrep_linear_bwd_x(g, w, ctx.plan) if nx else None, irrep_linear_bwd_w(x, g, ctx.plan) if nw else None, None)


irrep_linear.register_autograd(_lin_bwd, setup_context=_lin_setup)


def _lin_bx_setup(ctx, inputs, output):
    g, w, plan = inputs
    ctx.save_for_backward(g, w)
    ctx.plan = plan


def _lin_bx_bwd(ctx, v):  # dx = L^T(g; w);  <v, dx> = <L(v; w), g>
    g, w = ctx.saved_tensors
    ng, nw = ctx.needs_input_grad[:2]
    return (irrep_linear(v, w, ctx.plan).reshape(g.shape) if ng else None,
            irrep_linear_bwd_w(v, g, ctx.plan) if nw else None, None)


irrep_linear_bwd_x.register_autograd(_lin_bx_bwd, setup_context=_lin_bx_setup)


def _lin_bw_setup(ctx, inputs, output):
    x, g, plan = inputs
    ctx.save_for_backward(x, g)
    ctx.plan = plan


def _lin_bw_bwd(ctx, V):  # <V, dw> = <L(x; V), g>
    x, g = ctx.saved_tensors
    nx, ng = ctx.needs_input_grad[:2]
    return (irrep_linear_bwd_x(g, V, ctx.plan) if nx else None,
            irrep_linear(x, V, ctx.plan).reshape(g.shape) if ng else None, None)


irrep_linear_bwd_w.register_autograd(_lin_bw_bwd, setup_context=_lin_bw_setup)


# ======================================================================================================================
# element-indexed linear
# ======================================================================================================================
# PML_EL_SORTED=0: keep the FFMA kernels of element_linear.cu even when a SpeciesPlan is available (A/B runs)
EL_SORTED = [os.environ.get("PML_EL_SORTED", "1") == "1"]


def _el_species(pl, sp: int, n: int):
    """the SpeciesPlan behind handle `sp` if this call takes the tensor-core route (nodes grouped by element), else None"""
    if sp < 0 or not EL_SORTED[0] or not pl.seg_ok or GEMM_MODE[0] != "tc" or n < TC_MIN_ROWS[0]:
        return None
    s = _SPECIES.get(sp)
    if s is None:
        raise RuntimeError("the SpeciesPlan of this element_linear call was released before its gradient was taken")
    if s.N != n or s.Z != pl.Z:
        raise ValueError(f"SpeciesPlan of {s.N} nodes / {s.Z} elements used on {n} nodes / {pl.Z} elements")
    return s


def _el_gather(src: Tensor, src_row: int, s: "SpeciesPlan", colmap: Tensor, ncols: int) -> Tensor:
    """dense copy of `src` with the rows grouped by element and every block component-major"""
    dst = _new_empty(src, s.N, ncols)
    _call("pml_permute_rows_cols", _p(src), src_row, _p(s.order), _p(colmap), _p(dst), ncols, ncols, s.N, 0, _stream())
    return dst


def _el_tables(pl, s: "SpeciesPlan"):
    tabs = s.tables(pl)
    sz = C.sizeof(_lib.GemmProblem) * pl.seg_nprob
    _, _, cm_in, cm_out = pl.segment_tables(tabs.device)
    return tabs.data_ptr(), tabs.data_ptr() + sz, tabs.data_ptr() + 2 * sz, cm_in, cm_out


def _el_fwd_sorted(pl, s, x: Tensor, w: Tensor, xs: Optional[Tensor] = None) -> Tensor:
    """y = L(x; w) on nodes grouped by element: gather -> grouped tensor-core GEMM -> scatter"""
    n = x.shape[0]
    t_fwd, _, _, cm_in, cm_out = _el_tables(pl, s)
    if xs is None:
        xs = _el_gather(x, pl.d_in, s, cm_in, pl.d_in_c)
    ys = _new_empty(x, n, pl.d_out_c)
    _gemm("pml_gemm_grouped_tc", _p(xs), _p(w), _p(ys), t_fwd, pl.seg_nprob, n, n, pl.seg_max_mo, pl.seg_max_d, 1, pl.seg_max_mi)
    out = _zeros(x, _el_out_shape(pl, n)) if pl.has_unfed_output else _new_empty(x, _el_out_shape(pl, n))
    _call("pml_permute_rows_cols", _p(ys), pl.d_out_c, _p(s.order), _p(cm_out), _p(out), pl.d_out, pl.d_out_c, n, 1, _stream())
    return out


def _el_bx_sorted(pl, s, g: Tensor, w: Tensor, gs: Optional[Tensor] = None) -> Tensor:
    n = g.shape[0]
    _, t_bx, _, cm_in, cm_out = _el_tables(pl, s)
    if gs is None:
        gs = _el_gather(g, pl.d_out, s, cm_out, pl.d_out_c)
    dxs = _new_empty(g, n, pl.d_in_c)
    _gemm("pml_gemm_grouped_tc", _p(gs), _p(w), _p(dxs), t_bx, pl.seg_nprob, n, n, pl.seg_max_mi, pl.seg_max_d, 1, pl.seg_max_mo)
    dx = _zeros(g, n, pl.d_in) if pl.has_unfed_input else _new_empty(g, n, pl.d_in)
    _call("pml_permute_rows_cols", _p(dxs), pl.d_in_c, _p(s.order), _p(cm_in), _p(dx), pl.d_in, pl.d_in_c, n, 1, _stream())
    return dx


def _el_bw_sorted(pl, s, x: Tensor, g: Tensor, xs: Optional[Tensor] = None, gs: Optional[Tensor] = None) -> Tensor:
    n = x.shape[0]
    _, _, t_bw, cm_in, cm_out = _el_tables(pl, s)
    if xs is None:
        xs = _el_gather(x, pl.d_in, s, cm_in, pl.d_in_c)
    if gs is None:
        gs = _el_gather(g, pl.d_out, s, cm_out, pl.d_out_c)
    dw = _zeros(x, pl.weight_numel)
    # the reduction runs over one element's nodes (<= n): split so that a CTA sums <= 1008 terms on the plain TMEM
    # accumulator (see irrep_linear_bwd_w); the (2l+1) components of a block add into the same dW slab atomically
    sk = max(1, -(-n // 1008))
    _gemm("pml_gemm_grouped_tc", _p(xs), _p(gs), _p(dw), t_bw, pl.seg_nprob, n, pl.seg_max_mi, pl.seg_max_mo, pl.seg_max_d, sk,
          min((n + sk - 1) // sk + 16, 256))
    return dw


def _el_out_shape(pl, n: int):
    if pl.out_layout == "channel":
        S = sum(2 * l + 1 for _, l, _ in pl.irreps_out)
        return (n, pl.d_out // S, S)
    return (n, pl.d_out)


@torch.library.custom_op(f"{NS}::element_linear", mutates_args=(), device_types="cuda")
def element_linear(x: Tensor, attrs: Tensor, w: Tensor, plan: int, sp: int = -1) -> Tensor:
    pl = plans.get(plan)
    x, attrs, w = _f32c(x), _f32c(attrs), _f32c(w)
    n = x.shape[0]
    s = _el_species(pl, sp, n)
    if s is not None:
        return _el_fwd_sorted(pl, s, x, w)
    shape = _el_out_shape(pl, n)
    out = _zeros(x, shape) if pl.has_unfed_output else _new_empty(x, shape)
    _call("pml_element_linear_fwd", _p(x), _p(attrs), _p(w), C.byref(pl.c_plan), _p(out), n, pl.Z, 0, _stream(), launches=pl.n_launch)
    return out


@element_linear.register_fake
def _(x, attrs, w, plan, sp=-1):
    pl = plans.get(plan)
    if pl.out_layout == "channel":
        S = sum(2 * l + 1 for _, l, _ in pl.irreps_out)
        return _new_empty(x, x.shape[0], pl.d_out // S, S)
    return _new_empty(x, x.shape[0], pl.d_out)


@torch.library.custom_op(f"{NS}::element_linear_bwd_x", mutates_args=(), device_types="cuda")
def element_linear_bwd_x(g: Tensor, attrs: Tensor, w: Tensor, plan: int, sp: int = -1) -> Tensor:
    pl = plans.get(plan)
    g, attrs, w = _f32c(g), _f32c(attrs), _f32c(w)
    n = g.shape[0]
    s = _el_species(pl, sp, n)
    if s is not None:
        return _el_bx_sorted(pl, s, g, w)
    dx = _zeros(g, n, pl.d_in)  # accumulated atomically per (element, node chunk)
    _call("pml_element_linear_bwd_x", _p(g), _p(attrs), _p(w), C.byref(pl.c_plan), _p(dx), n, pl.Z, 0, _stream(), launches=pl.n_launch)
    return dx


@element_linear_bwd_x.register_fake
def _(g, attrs, w, plan, sp=-1):
    return _new_empty(g, g.shape[0], plans.get(plan).d_in)


@torch.library.custom_op(f"{NS}::element_linear_bwd_w", mutates_args=(), device_types="cuda")
def element_linear_bwd_w(x: Tensor, g: Tensor, attrs: Tensor, plan: int, sp: int = -1) -> Tensor:
    pl = plans.get(plan)
    x, g, attrs = _f32c(x), _f32c(g), _f32c(attrs)
    s = _el_species(pl, sp, x.shape[0])
    if s is not None:
        return _el_bw_sorted(pl, s, x, g)
    dw = _zeros(x, pl.weight_numel)
    _call("pml_element_linear_bwd_w", _p(x), _p(g), _p(attrs), C.byref(pl.c_plan), _p(dw), x.shape[0], pl.Z, 0, _stream(),
          launches=pl.n_launch)
    return dw


@element_linear_bwd_w.register_fake
def _(x, g, attrs, plan, sp=-1):
    return _new_empty(x, plans.get(plan).weight_numel)


def _el_setup(ctx, inputs, output):
    x, attrs, w, plan, sp = inputs
    ctx.save_for_backward(x, attrs, w)
    ctx.plan, ctx.sp = plan, sp
    ctx.sp_ref = _SPECIES.get(sp) if sp >= 0 else None  # keeps the plan alive as long as the autograd graph


def _el_shared(ctx, n: int):
    """SpeciesPlan for a cotangent formula that may gather an operand ONCE for the two results that read it (only when no
    graph is being recorded: the results then need no autograd node of their own)"""
    if torch.is_grad_enabled():
        return None, None
    pl = plans.get(ctx.plan)
    return pl, _el_species(pl, ctx.sp, n)


def _el_bwd(ctx, g):
    x, attrs, w = ctx.saved_tensors
    nx, _, nw = ctx.needs_input_grad[:3]
    nw = nw and _PARAM_GRADS[0]
    pl, s = _el_shared(ctx, x.shape[0]) if (nx and nw) else (None, None)
    if s is not None:  # dx and dW both read the gathered cotangent
        g_, x_, w_ = _f32c(g), _f32c(x), _f32c(w)
        gs = _el_gather(g_, pl.d_out, s, _el_tables(pl, s)[4], pl.d_out_c)
        return _el_bx_sorted(pl, s, g_, w_, gs), None, _el_bw_sorted(pl, s, x_, g_, None, gs), None, None
    return (element_linear_bwd_x(g, attrs, w, ctx.plan, ctx.sp) if nx else None, None,
            element_linear_bwd_w(x, g, attrs, ctx.plan, ctx.sp) if nw else None, None, None)


element_linear.register_autograd(_el_bwd, setup_context=_el_setup)


def _el_bx_setup(ctx, inputs, output):
    g, attrs, w, plan, sp = inputs
    ctx.save_for_backward(g, attrs, w)
    ctx.plan, ctx.sp = plan, sp
    ctx.sp_ref = _SPECIES.get(sp) if sp >= 0 else None


def _el_bx_bwd(ctx, v):
    g, attrs, w = ctx.saved_tensors
    ng, _, nw = ctx.needs_input_grad[:3]
    pl, s = _el_shared(ctx, g.shape[0]) if (ng and nw) else (None, None)
    if s is not None:  # L(v; w) and dW(v, g) both read the gathered tangent
        v_, g_, w_ = _f32c(v), _f32c(g), _f32c(w)
        vs = _el_gather(v_, pl.d_in, s, _el_tables(pl, s)[3], pl.d_in_c)
        return _el_fwd_sorted(pl, s, v_, w_, vs).reshape(g.shape), None, _el_bw_sorted(pl, s, v_, g_, vs, None), None, None
    return (element_linear(v, attrs, w, ctx.plan, ctx.sp).reshape(g.shape) if ng else None, None,
            element_linear_bwd_w(v, g, attrs, ctx.plan, ctx.sp) if nw else None, None, None)


element_linear_bwd_x.register_autograd(_el_bx_bwd, setup_context=_el_bx_setup)


def _el_bw_setup(ctx, inputs, output):
    x, g, attrs, plan, sp = inputs
    ctx.save_for_backward(x, g, attrs)
    ctx.plan, ctx.sp = plan, sp
    ctx.sp_ref = _SPECIES.get(sp) if sp >= 0 else None


def _el_bw_bwd(ctx, V):
    x, g, attrs = ctx.saved_tensors
    nx, ng = ctx.needs_input_grad[:2]
    return (element_linear_bwd_x(g, attrs, V, ctx.plan, ctx.sp) if nx else None,
            element_linear(x, attrs, V, ctx.plan, ctx.sp).reshape(g.shape) if ng else None, None, None, None)


element_linear_bwd_w.register_autograd(_el_bw_bwd, setup_context=_el_bw_setup)


# ======================================================================================================================
# symmetric contraction
# ======================================================================================================================
@torch.library.custom_op(f"{NS}::sym_contract", mutates_args=(), device_types="cuda")
def sym_contract(x: Tensor, attrs: Tensor, weights: List[Tensor], plan: int) -> Tensor:
    pl = plans.get(plan)
    x, attrs = _f32c(x), _f32c(attrs)
    ws = [_f32c(w) for w in weights]
    n = x.shape[0]
    out = _new_empty(x, n, pl.C * pl.DOUT)
    st = pl.weight_struct(ws)
    _call("pml_sym_fwd", pl.spec_id, _p(x), _p(attrs), C.byref(st), _p(out), n, pl.C, pl.Z, _stream())
    return out


@sym_contract.register_fake
def _(x, attrs, weights, plan):
    pl = plans.get(plan)
    return _new_empty(x, x.shape[0], pl.C * pl.DOUT)


@torch.library.custom_op(f"{NS}::sym_contract_bwd", mutates_args=(), device_types="cuda")
def sym_contract_bwd(x: Tensor, attrs: Tensor, weights: List[Tensor], g: Tensor, plan: int, need_x: bool,
                     need_w: bool) -> List[Tensor]:
    """returns [dx, dW_0, dW_1, ...] (empty tensors for what was not requested)"""
    pl = plans.get(plan)
    x, attrs, g = _f32c(x), _f32c(attrs), _f32c(g)
    ws = [_f32c(w) for w in weights]
    n = x.shape[0]
    dx = _empty_like(x) if need_x else _new_empty(x, 0)
    gws = [_zeros_like(w) for w in ws] if need_w else None
    st = pl.weight_struct(ws, gws)
    if need_x or need_w:
        _call("pml_sym_bwd", pl.spec_id, _p(x), _p(attrs), C.byref(st), _p(g), _p(dx) if need_x else None, n, pl.C, pl.Z,
              _stream())
    return [dx] + (gws if need_w else [_new_empty(x, 0) for _ in ws])


@sym_contract_bwd.register_fake
def _(x, attrs, weights, g, plan, need_x, need_w):
    return [_empty_like(x) if need_x else _new_empty(x, 0)] + [
        _empty_like(w) if need_w else _new_empty(x, 0) for w in weights
    ]


@torch.library.custom_op(f"{NS}::sym_contract_dbl", mutates_args=(), device_types="cuda")
def sym_contract_dbl(x: Tensor, v: Tensor, attrs: Tensor, weights: List[Tensor], g: Tensor, plan: int, need_g: bool,
                     need_x: bool, need_w: bool) -> List[Tensor]:
    """gradient of d/de <g, out(x + e v, W)> w.r.t. (g, x, W): returns [dg, ddx, ddW_0, ...]"""
    pl = plans.get(plan)
    x, v, attrs, g = _f32c(x), _f32c(v), _f32c(attrs), _f32c(g)
    ws = [_f32c(w) for w in weights]
    n = x.shape[0]
    dg = _empty_like(g) if need_g else _new_empty(x, 0)
    ddx = _empty_like(x) if need_x else _new_empty(x, 0)
    gws = [_zeros_like(w) for w in ws] if need_w else None
    st = pl.weight_struct(ws, gws)
    if need_g or need_x or need_w:
        _call("pml_sym_dbl", pl.spec_id, _p(x), _p(v), _p(attrs), C.byref(st), _p(g), _p(dg) if need_g else None,
              _p(ddx) if need_x else None, n, pl.C, pl.Z, _stream())
    return [dg, ddx] + (gws if need_w else [_new_empty(x, 0) for _ in ws])


@sym_contract_dbl.register_fake
def _(x, v, attrs, weights, g, plan, need_g, need_x, need_w):
    return [_empty_like(g) if need_g else _new_empty(x, 0), _empty_like(x) if need_x else _new_empty(x, 0)] + [
        _empty_like(w) if need_w else _new_empty(x, 0) for w in weights
    ]


def _sym_setup(ctx, inputs, output):
    x, attrs, weights, plan = inputs
    ctx.save_for_backward(x, attrs, *weights)
    ctx.plan = plan


def _sym_bwd(ctx, g):
    x, attrs, *ws = ctx.saved_tensors
    nx = ctx.needs_input_grad[0]
    nw = ctx.needs_input_grad[2]
    nw_any = (any(nw) if isinstance(nw, (list, tuple)) else bool(nw)) and _PARAM_GRADS[0]
    res = sym_contract_bwd(x, attrs, list(ws), g, ctx.plan, nx, nw_any)
    return (res[0] if nx else None, None, list(res[1:]) if nw_any else [None] * len(ws), None)


sym_contract.register_autograd(_sym_bwd, setup_context=_sym_setup)


def _sym_bwd_setup(ctx, inputs, output):
    x, attrs, weights, g, plan, need_x, need_w = inputs
    ctx.set_materialize_grads(False)
    ctx.save_for_backward(x, attrs, g, *weights)
    ctx.plan = plan
    ctx.needs = (need_x, need_w)


def _sym_bwd_bwd(ctx, grads):
    # grads: list aligned with the op's output list [dx, dW...]
    x, attrs, g, *ws = ctx.saved_tensors
    nx, _, nw, ng = ctx.needs_input_grad[:4]
    nw_any = any(nw) if isinstance(nw, (list, tuple)) else bool(nw)
    vx = grads[0] if ctx.needs[0] else None
    vws = grads[1:] if ctx.needs[1] else [None] * len(ws)
    gx = gg = None
    gw = [None] * len(ws)
    if vx is not None:
        r = sym_contract_dbl(x, vx, attrs, list(ws), g, ctx.plan, ng, nx, nw_any)
        gg = r[0] if ng else None
        gx = r[1] if nx else None
        if nw_any:
            gw = list(r[2:])
    if any(v is not None for v in vws):
        # Phi is linear in W: <vW, dPhi/dW> = Phi(x, vW, g)
        vw = [v if v is not None else torch.zeros_like(w) for v, w in zip(vws, ws)]
        if ng:
            gg = _acc(gg, sym_contract(x, attrs, vw, ctx.plan))
        if nx:
            gx = _acc(gx, sym_contract_bwd(x, attrs, vw, g, ctx.plan, True, False)[0])
    return (gx, None, gw if nw_any else [None] * len(ws), gg, None, None, None)


sym_contract_bwd.register_autograd(_sym_bwd_bwd, setup_context=_sym_bwd_setup)


# ======================================================================================================================
# activations, pooling, layout
# ======================================================================================================================
@torch.library.custom_op(f"{NS}::act", mutates_args=(), device_types="cuda")
def act(x: Tensor, kind: int, c: float) -> Tensor:
    x = _f32c(x)
    y = _empty_like(x)
    _call("pml_act", _p(x), None, None, float(c), kind, 0, _p(y), x.numel(), _stream())
    return y


@act.register_fake
def _(x, kind, c):
    return _empty_like(x)


@torch.library.custom_op(f"{NS}::act_bwd", mutates_args=(), device_types="cuda")
def act_bwd(g: Tensor, x: Tensor, kind: int, c: float) -> Tensor:
    g, x = _f32c(g), _f32c(x)
    y = _empty_like(x)
    _call("pml_act", _p(x), _p(g), None, float(c), kind, 1, _p(y), x.numel(), _stream())
    return y


@act_bwd.register_fake
def _(g, x, kind, c):
    return _empty_like(x)


@torch.library.custom_op(f"{NS}::act_bwd2", mutates_args=(), device_types="cuda")
def act_bwd2(g: Tensor, v: Tensor, x: Tensor, kind: int, c: float) -> Tensor:
    g, v, x = _f32c(g), _f32c(v), _f32c(x)
    y = _empty_like(x)
    _call("pml_act", _p(x), _p(g), _p(v), float(c), kind, 2, _p(y), x.numel(), _stream())
    return y


@act_bwd2.register_fake
def _(g, v, x, kind, c):
    return _empty_like(x)


def _act_setup(ctx, inputs, output):
    x, kind, c = inputs
    ctx.save_for_backward(x)
    ctx.kc = (kind, c)


act.register_autograd(lambda ctx, g: (act_bwd(g, ctx.saved_tensors[0], *ctx.kc), None, None), setup_context=_act_setup)


def _act_bwd_setup(ctx, inputs, output):
    g, x, kind, c = inputs
    ctx.save_for_backward(g, x)
    ctx.kc = (kind, c)


def _act_bwd_bwd(ctx, v):
    g, x = ctx.saved_tensors
    ng, nx = ctx.needs_input_grad[:2]
    return (act_bwd(v, x, *ctx.kc) if ng else None, act_bwd2(g, v, x, *ctx.kc) if nx else None, None, None)


act_bwd.register_autograd(_act_bwd_bwd, setup_context=_act_bwd_setup)


# NequIP gate (reference nequip/modules/_gate.py:131-152): value, cotangent map, derivative of the cotangent map
@torch.library.custom_op(f"{NS}::gate", mutates_args=(), device_types="cuda")
def gate(x: Tensor, plan: int) -> Tensor:
    pl = plans.get(plan)
    x = _f32c(x)
    y = _new_empty(x, x.shape[0], pl.d_out)
    _call("pml_gate", _p(x), None, None, C.byref(pl.c_plan), 0, _p(y), None, x.shape[0], _stream())
    return y


@gate.register_fake
def _(x, plan):
    return _new_empty(x, x.shape[0], plans.get(plan).d_out)


@torch.library.custom_op(f"{NS}::gate_bwd", mutates_args=(), device_types="cuda")
def gate_bwd(x: Tensor, dy: Tensor, plan: int) -> Tensor:
    pl = plans.get(plan)
    x, dy = _f32c(x), _f32c(dy)
    dx = _empty_like(x)
    _call("pml_gate", _p(x), _p(dy), None, C.byref(pl.c_plan), 1, _p(dx), None, x.shape[0], _stream())
    return dx


@gate_bwd.register_fake
def _(x, dy, plan):
    return _empty_like(x)


@torch.library.custom_op(f"{NS}::gate_bwd2", mutates_args=(), device_types="cuda")
def gate_bwd2(x: Tensor, dy: Tensor, v: Tensor, plan: int) -> Tuple[Tensor, Tensor]:
    """(ddy, ddx): gradients of <v, gate_bwd(x, dy)> w.r.t. dy and x"""
    pl = plans.get(plan)
    x, dy, v = _f32c(x), _f32c(dy), _f32c(v)
    ddy, ddx = _empty_like(dy), _empty_like(x)
    _call("pml_gate", _p(x), _p(dy), _p(v), C.byref(pl.c_plan), 2, _p(ddy), _p(ddx), x.shape[0], _stream())
    return ddy, ddx


@gate_bwd2.register_fake
def _(x, dy, v, plan):
    return _empty_like(dy), _empty_like(x)


def _gate_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[0])
    ctx.plan = inputs[1]


gate.register_autograd(lambda ctx, g: (gate_bwd(ctx.saved_tensors[0], g, ctx.plan), None), setup_context=_gate_setup)


def _gate_bwd_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[0], inputs[1])
    ctx.plan = inputs[2]


def _gate_bwd_bwd(ctx, v):
    x, dy = ctx.saved_tensors
    nx, ng = ctx.needs_input_grad[:2]
    ddy, ddx = gate_bwd2(x, dy, v, ctx.plan)
    return (ddx if nx else None, ddy if ng else None, None)


gate_bwd.register_autograd(_gate_bwd_bwd, setup_context=_gate_bwd_setup)


@torch.library.custom_op(f"{NS}::pool_sum", mutates_args=(), device_types="cuda")
def pool_sum(x: Tensor, batch: Tensor, num_graphs: int) -> Tensor:
    x = _f32c(x)
    out = _zeros(x, num_graphs, x.shape[1])
    _call("pml_pool_sum", _p(x), _p(batch), _p(out), x.shape[0], x.shape[1], _stream())
    return out


@pool_sum.register_fake
def _(x, batch, num_graphs):
    return _new_empty(x, num_graphs, x.shape[1])


@torch.library.custom_op(f"{NS}::pool_gather", mutates_args=(), device_types="cuda")
def pool_gather(g: Tensor, batch: Tensor) -> Tensor:
    g = _f32c(g)
    out = _new_empty(g, batch.numel(), g.shape[1])
    _call("pml_pool_gather", _p(g), _p(batch), _p(out), batch.numel(), g.shape[1], _stream())
    return out


@pool_gather.register_fake
def _(g, batch):
    return _new_empty(g, batch.numel(), g.shape[1])


def _pool_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[1])
    ctx.G = inputs[2]


pool_sum.register_autograd(lambda ctx, g: (pool_gather(g, ctx.saved_tensors[0]), None, None), setup_context=_pool_setup)


def _poolg_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[1])
    ctx.G = inputs[0].shape[0]


pool_gather.register_autograd(lambda ctx, v: (pool_sum(v, ctx.saved_tensors[0], ctx.G), None), setup_context=_poolg_setup)


@torch.library.custom_op(f"{NS}::irreps_to_channel", mutates_args=(), device_types="cuda")
def irreps_to_channel(x: Tensor, channels: int, lmax: int, inverse: bool) -> Tensor:
    """[N, C*S] irreps layout -> [N, C, S] (inverse=False) or back (inverse=True); irreps_tools.py:47-67."""
    x = _f32c(x)
    n, S = x.shape[0], (lmax + 1) ** 2
    y = _new_empty(x, n, channels * S) if inverse else _new_empty(x, n, channels, S)
    _call("pml_irreps_to_channel", _p(x), _p(y), n, channels, lmax, 1 if inverse else 0, _stream())
    return y


@irreps_to_channel.register_fake
def _(x, channels, lmax, inverse):
    S = (lmax + 1) ** 2
    return _new_empty(x, x.shape[0], channels * S) if inverse else _new_empty(x, x.shape[0], channels, S)


def _i2c_setup(ctx, inputs, output):
    ctx.a = inputs[1:]


irreps_to_channel.register_autograd(
    lambda ctx, g: (irreps_to_channel(g, ctx.a[0], ctx.a[1], not ctx.a[2]), None, None, None), setup_context=_i2c_setup
)


# ======================================================================================================================
# fused force-matching loss
# ======================================================================================================================
@torch.library.custom_op(f"{NS}::energy_force_mse", mutates_args=(), device_types="cuda")
def energy_force_mse(e: Tensor, e_t: Tensor, gpos: Tensor, f_t: Tensor, w_e: float, w_f: float) -> Tensor:
    """w_e * mean((e - e_t)^2) + w_f * mean((-gpos - f_t)^2) with forces = -gpos = -dE/dx (reference lightning/module.py:65-95
    + the stock MSE losses on y_graph_scalars / y_node_vector, mace_module.py:159-185): one deterministic kernel"""
    e, e_t, gpos, f_t = _f32c(e), _f32c(e_t), _f32c(gpos), _f32c(f_t)
    if e.numel() != e_t.numel() or gpos.numel() != f_t.numel():
        raise ValueError("prediction / target shapes differ")
    out = _new_empty(e, 1)
    _call("pml_ef_mse_fwd", _p(e), _p(e_t), e.numel(), _p(gpos), _p(f_t), gpos.numel(), float(w_e), float(w_f), _p(out), _stream())
    return out.reshape(())


@energy_force_mse.register_fake
def _(e, e_t, gpos, f_t, w_e, w_f):
    return e.new_empty(())


@torch.library.custom_op(f"{NS}::energy_force_mse_bwd", mutates_args=(), device_types="cuda")
def energy_force_mse_bwd(gl: Tensor, e: Tensor, e_t: Tensor, gpos: Tensor, f_t: Tensor, w_e: float, w_f: float) -> Tuple[Tensor, Tensor]:
    e, e_t, gpos, f_t, gl = _f32c(e), _f32c(e_t), _f32c(gpos), _f32c(f_t), _f32c(gl)
    de, dg = _empty_like(e), _empty_like(gpos)
    _call("pml_ef_mse_bwd", _p(e), _p(e_t), e.numel(), _p(gpos), _p(f_t), gpos.numel(), float(w_e), float(w_f), _p(gl), _p(de), _p(dg),
          _stream())
    return de, dg


@energy_force_mse_bwd.register_fake
def _(gl, e, e_t, gpos, f_t, w_e, w_f):
    return _empty_like(e), _empty_like(gpos)


def _efm_setup(ctx, inputs, output):
    e, e_t, gpos, f_t, w_e, w_f = inputs
    ctx.save_for_backward(e, e_t, gpos, f_t)
    ctx.w = (w_e, w_f)


def _efm_bwd(ctx, gl):
    e, e_t, gpos, f_t = ctx.saved_tensors
    de, dg = energy_force_mse_bwd(gl, e, e_t, gpos, f_t, *ctx.w)
    return (de if ctx.needs_input_grad[0] else None, None, dg if ctx.needs_input_grad[2] else None, None, None, None)


energy_force_mse.register_autograd(_efm_bwd, setup_context=_efm_setup)


# ======================================================================================================================
# graph structure (index plumbing only — no floating-point work)
# ======================================================================================================================
_SPECIES: "weakref.WeakValueDictionary[int, SpeciesPlan]" = weakref.WeakValueDictionary()
_SPECIES_NEXT = [0]
_ONE_HOT_SEEN: list = []  # (data_ptr, version, shape) of attrs tensors already validated (most recent last)


class SpeciesPlan:
    """The nodes of a batch grouped by chemical element (``node_attrs`` one-hot rows): ``order`` [N] node ids, element by
    element, and ``seg`` [Z + 2] group starts (+ a not-one-hot flag), both on the device and built by ONE kernel with no
    host round trip — so the plan can be rebuilt inside a captured CUDA graph from whatever batch sits in the static
    buffers.  It turns the element-indexed linears (mix_layer / residual_connection_layer / self_connection_layer) into
    grouped tensor-core GEMMs over contiguous row ranges; ``tables(plan)`` are the per-batch problem tables of one
    ElementLinearPlan (three: forward, dx, dW), also completed on the device."""

    def __init__(self, attrs: Tensor) -> None:
        attrs = _f32c(attrs)
        self.N, self.Z = int(attrs.shape[0]), int(attrs.shape[1])
        dev = attrs.device
        self.order = torch.empty(max(self.N, 1), dtype=torch.int32, device=dev)
        self.seg = torch.empty(self.Z + 2, dtype=torch.int32, device=dev)
        _call("pml_species_order", _p(attrs), self.N, self.Z, _p(self.order), _p(self.seg), _stream())
        self._src = (attrs.data_ptr(), attrs._version, tuple(attrs.shape))
        self._tables: dict = {}
        self.handle = _SPECIES_NEXT[0]
        _SPECIES_NEXT[0] += 1
        _SPECIES[self.handle] = self

    def matches(self, attrs: Tensor) -> bool:
        return self._src == (attrs.data_ptr(), attrs._version, tuple(attrs.shape))

    def one_hot(self) -> bool:
        """host check (one synchronising read of the flag the kernel wrote)"""
        return int(self.seg[self.Z + 1].item()) == 0

    def tables(self, pl) -> Tensor:
        t = self._tables.get(pl.handle)
        if t is None:
            tmpl, meta, _, _ = pl.segment_tables(self.order.device)
            n = 3 * pl.seg_nprob
            t = torch.empty(n * C.sizeof(_lib.GemmProblem), dtype=torch.uint8, device=self.order.device)
            _call("pml_segment_problems", _p(tmpl), _p(meta), n, _p(self.seg), self.Z, _p(t), _stream())
            self._tables[pl.handle] = t
        return t


def species_plan_for(data: dict, el_plans=()) -> Optional["SpeciesPlan"]:
    """SpeciesPlan of ``data["node_attrs"]`` (the caller-supplied ``data["_species_plan"]`` if it still describes that
    tensor), or None when the FFMA kernels should run: attrs not one-hot, more than 128 elements, too few nodes to fill
    128-row MMA tiles, or no element-indexed linear large enough to profit.  ``el_plans``: the ElementLinearPlans whose
    problem tables are built right away (on the caller's stream, before any side-stream work uses them).

    One-hot rows are validated with one synchronising read the first time an attrs tensor is seen; inside a CUDA-graph
    capture the validation of the warm-up passes stands, and a later batch that is not one-hot turns the outputs into
    NaN (pml_segment_problems) instead of silently using the wrong weights."""
    attrs = data["node_attrs"]
    if not (EL_SORTED[0] and GEMM_MODE[0] == "tc" and attrs.is_cuda and attrs.dim() == 2):
        return None
    n, z = attrs.shape
    el_plans = [pl for pl in el_plans if pl.seg_ok]
    if n < TC_MIN_ROWS[0] or z > 128 or not el_plans:
        return None
    sp = data.get("_species_plan")
    if sp is None or not sp.matches(attrs):
        sp = SpeciesPlan(attrs)
        if not torch.cuda.is_current_stream_capturing():
            key = sp._src
            if key in _ONE_HOT_SEEN:
                _ONE_HOT_SEEN.remove(key)
            elif not sp.one_hot():
                return None
            _ONE_HOT_SEEN.append(key)
            del _ONE_HOT_SEEN[:-16]
    for pl in el_plans:
        sp.tables(pl)
    return sp


def edge_plan_for(data: dict, num_nodes: int, scatter_row: int) -> "EdgePlan":
    """the caller-supplied ``data["_edge_plan"]`` (e.g. from ``neighbour_list.radius_graph``) if it still describes
    ``data["edge_index"]``, else a fresh plan; never written back into the caller's dict"""
    plan = data.get("_edge_plan")
    if plan is not None and plan.matches(data["edge_index"], num_nodes):
        return plan
    return EdgePlan(data["edge_index"], num_nodes, scatter_row=scatter_row)


class EdgePlan:
    """int32 copies of the edge list and the CSR over the scatter index, built once per batch and shared by every
    layer's forward and backward kernels.  ``gather`` / ``rowptr`` / ``perm`` are the view for MACE's convention
    (gather at edge_index[0], scatter to edge_index[1]); ``view(0)`` gives NequIP's (gather [1], scatter [0])."""

    def __init__(self, edge_index: Tensor, num_nodes: int, scatter_row: int = 1, assume_grouped: bool = False) -> None:
        E = edge_index.shape[1]
        dev = edge_index.device
        ei = edge_index.contiguous()
        self.E, self.N = E, num_nodes
        self.snd = torch.empty(E, dtype=torch.int32, device=dev)
        self.rcv = torch.empty(E, dtype=torch.int32, device=dev)
        _call("pml_split_edge_index", _p(ei), _p(self.snd), _p(self.rcv), E, _stream())
        self._views = {}
        self._ei = ei
        self._src = (edge_index.data_ptr(), edge_index._version, tuple(edge_index.shape))
        self._assume_grouped = assume_grouped
        self.scatter_row = scatter_row
        self.gather, self.rowptr, self.perm = self.view(scatter_row)

    @classmethod
    def from_sorted(cls, snd: Tensor, rcv: Tensor, rowptr: Tensor, perm_by_receiver: Tensor, num_nodes: int) -> "EdgePlan":
        """From the neighbour-list kernel: edges sorted by sender, symmetric set => one row pointer serves both views."""
        self = cls.__new__(cls)
        self.E, self.N = snd.numel(), num_nodes
        self.snd, self.rcv = snd, rcv
        self._views = {1: (snd, rowptr, perm_by_receiver), 0: (rcv, rowptr, None)}
        self._ei = None
        self._src = None  # set by the neighbour list to the edge_index tensor it returned alongside
        self._assume_grouped = False
        self.scatter_row = 1
        self.gather, self.rowptr, self.perm = self._views[1]
        return self

    def matches(self, edge_index: Tensor, num_nodes: int) -> bool:
        """False when this plan was derived from a different (or since modified) edge_index tensor: a caller that
        reuses a batch dict after a neighbour-list rebuild must not silently get the old edges"""
        if self.N != num_nodes or self.E != edge_index.shape[1]:
            return False
        if self._src is None:
            return True
        return self._src == (edge_index.data_ptr(), edge_index._version, tuple(edge_index.shape))

    def view(self, scatter_row: int):
        """(gather index, CSR row pointer over the scatter index, edge ids grouped by scatter index or None)"""
        if scatter_row not in self._views:
            ei, dev = self._ei, self._ei.device
            scratch = torch.zeros(2, self.N, dtype=torch.int32, device=dev)  # counts, cursor
            rowptr = torch.empty(self.N + 1, dtype=torch.int32, device=dev)
            _call("pml_csr_rowptr", _p(ei[scatter_row]), _p(scratch[0]), _p(rowptr), self.E, self.N, _stream(), launches=2)
            perm = None
            if not self._assume_grouped:  # stable counting sort by the scatter index (own kernels: graph-capturable)
                perm = torch.empty(self.E, dtype=torch.int32, device=dev)
                _call("pml_csr_group", _p(ei[scatter_row]), _p(rowptr), _p(scratch[1]), _p(perm), self.E, self.N, _stream(),
                      launches=2)
            gather = self.snd if scatter_row == 1 else self.rcv
            self._views[scatter_row] = (gather, rowptr, perm)
        return self._views[scatter_row]
