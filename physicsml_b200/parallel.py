"""Data-parallel plumbing: one process per GPU, molecule batches sharded by rank, gradients averaged once per step
(the reference delegates this to Lightning DDP, lightning/model.py:159-163).

:class:`GradientReducer` packs the gradients of a FIXED parameter list into a few flat buckets (one pack kernel per
bucket, no ``torch.cat`` / per-tensor ``copy_``), and starts each bucket's NCCL all-reduce from a post-accumulate-grad
hook the moment the bucket's last gradient exists, so the reduction of the late layers overlaps the backward pass of
the early ones (NCCL runs on its own stream; the wait is issued in :meth:`finish`).  The optimiser reads the reduced
gradients straight from the flat buffers (``attach``), with the 1/world scale folded into the fused Adam kernel.

Parameters without a gradient on some rank (an element absent from that rank's batch) are packed as zeros, so every rank
reduces the same layout whatever its batch contains.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch
import torch.distributed as dist

CHUNK = 4096


class _PackTensor(C.Structure):
    _fields_ = [("src", C.c_void_p), ("offset", C.c_longlong), ("n", C.c_longlong)]


class _Chunk(C.Structure):
    _fields_ = [("tensor", C.c_int), ("pad", C.c_int), ("offset", C.c_longlong)]


def _world(group=None) -> int:
    return dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1


class GradientReducer:
    def __init__(self, params, n_buckets: int = 2, group=None, average: bool = False) -> None:
        """params: the parameters to reduce, in an order that is identical on every rank.  ``average=False`` leaves the
        SUM in the flat buffers (the fused Adam kernel applies 1/world as its ``grad_scale``); True divides in place."""
        self.params: List[torch.nn.Parameter] = [p for p in params if p.numel() > 0]
        self.group, self.average = group, average
        self.world = _world(group)
        self.n_buckets = max(1, min(n_buckets, len(self.params)))
        self.order: Optional[List[int]] = None  # parameter indices in the order their gradients complete in backward
        self._observed: List[int] = []
        self._hooks = [p.register_post_accumulate_grad_hook(self._make_hook(i)) for i, p in enumerate(self.params)]
        self._built = False
        self._works: list = []
        self._keep: list = []
        self.launched_early = 0  # buckets whose all-reduce started from a hook (= overlapped with backward)

    # ---- layout ------------------------------------------------------------------------------------------------------
    def _build(self) -> None:
        """bucket assignment from the completion order observed in the first backward (rank 0's, broadcast so that all
        ranks agree): the first half of the gradient bytes to complete forms bucket 0, etc."""
        n = len(self.params)
        order = self._observed + [i for i in range(n) if i not in set(self._observed)]
        if self.world > 1:
            t = torch.tensor(order, dtype=torch.int64, device=self.params[0].device if dist.get_backend(self.group) == "nccl" else "cpu")
            dist.broadcast(t, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)
            order = [int(x) for x in t.cpu()]
        self.order = order
        total = sum(self.params[i].numel() for i in order)
        self.bucket_of = [0] * n
        self.offset_of = [0] * n
        sizes = [0] * self.n_buckets
        acc, b = 0, 0
        for i in order:
            if b < self.n_buckets - 1 and acc >= (b + 1) * total / self.n_buckets:
                b += 1
            self.bucket_of[i], self.offset_of[i] = b, sizes[b]
            sizes[b] += (self.params[i].numel() + 3) // 4 * 4  # 16-byte aligned segments
            acc += self.params[i].numel()
        dev = self.params[0].device
        self.flat = [torch.zeros(max(s, 4), dtype=torch.float32, device=dev) for s in sizes]
        self.members = [[i for i in range(n) if self.bucket_of[i] == b] for b in range(self.n_buckets)]
        self.views = {}
        for i, p in enumerate(self.params):
            b, off = self.bucket_of[i], self.offset_of[i]
            self.views[p] = self.flat[b][off : off + p.numel()].view_as(p)
        if dev.type == "cuda":
            self._tables = []
            for b in range(self.n_buckets):
                chunks = []
                for k, i in enumerate(self.members[b]):
                    chunks += [(k, off) for off in range(0, self.params[i].numel(), CHUNK)]
                cs = (_Chunk * max(1, len(chunks)))(*[_Chunk(k, 0, off) for k, off in chunks])
                nt = len(self.members[b])
                tbytes = C.sizeof(_PackTensor) * nt
                pad = (-tbytes) % 16
                host = torch.zeros(tbytes + pad + C.sizeof(_Chunk) * len(chunks), dtype=torch.uint8).pin_memory()
                host[tbytes + pad :] = torch.frombuffer(bytearray(bytes(cs)), dtype=torch.uint8)[: C.sizeof(_Chunk) * len(chunks)]
                # a second pinned copy for a CUDA-graph capture: the captured memcpy node re-reads its source at every
                # replay, so it must not be the buffer eager steps keep rewriting (allocated now: no pinning inside a capture)
                self._tables.append({"host": host, "host_capture": host.clone().pin_memory(), "dev": torch.empty_like(host, device=dev),
                                     "chunk_off": tbytes + pad, "nchunks": len(chunks), "sig": None, "event": None})
        self._pending = [len(m) for m in self.members]
        self._done = [False] * self.n_buckets
        self._built = True
        if getattr(self, "_opt", None) is not None:
            self._opt.grad_views = self.views

    def attach(self, optimizer) -> None:
        """let a FusedAdam read the reduced gradients straight from the flat buckets (no unpack)"""
        self._opt = optimizer
        optimizer.grad_views = self.views if self._built else {}

    # ---- per step ----------------------------------------------------------------------------------------------------
    def _make_hook(self, i: int):
        def hook(_p):
            if not self._built:
                self._observed.append(i)
                return
            b = self.bucket_of[i]
            self._pending[b] -= 1
            if self._pending[b] == 0 and not self._done[b]:
                self.launched_early += 1
                self._launch(b)

        return hook

    def _pack(self, b: int) -> None:
        flat = self.flat[b]
        if flat.is_cuda:
            from . import ops

            tab = self._tables[b]
            ptrs = tuple((self.params[i].grad.data_ptr() if self.params[i].grad is not None else 0) for i in self.members[b])
            if ptrs != tab["sig"]:
                capturing = torch.cuda.is_current_stream_capturing()
                if tab["event"] is not None and not capturing:
                    tab["event"].synchronize()  # the previous upload has left the pinned buffer
                src = tab["host_capture"] if capturing else tab["host"]
                ts = (_PackTensor * len(ptrs))(*[
                    _PackTensor(ptr or None, self.offset_of[i], self.params[i].numel()) for ptr, i in zip(ptrs, self.members[b])])
                raw = torch.frombuffer(bytearray(bytes(ts)), dtype=torch.uint8)
                src[: raw.numel()] = raw
                tab["dev"].copy_(src, non_blocking=True)
                if not capturing:
                    tab["event"] = torch.cuda.Event()
                    tab["event"].record()
                tab["sig"] = ptrs
            base = tab["dev"].data_ptr()
            ops._call("pml_pack_f32", base, base + tab["chunk_off"], tab["nchunks"], flat.data_ptr(), ops._stream())
        else:  # host tensors (gloo tests of the plumbing): same layout, plain copies
            for i in self.members[b]:
                p, off = self.params[i], self.offset_of[i]
                seg = flat[off : off + p.numel()]
                if p.grad is None:
                    seg.zero_()
                else:
                    seg.copy_(p.grad.reshape(-1))

    def _launch(self, b: int) -> None:
        self._done[b] = True
        self._pack(b)
        if self.world > 1:
            self._works.append(dist.all_reduce(self.flat[b], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self) -> None:
        """after backward: start what no hook started (first step; parameters without gradient), wait for all buckets"""
        if not self._built:
            self._build()
        for b in range(self.n_buckets):
            if not self._done[b]:
                self._launch(b)
        for w in self._works:
            w.wait()
        self._works = []
        if self.average and self.world > 1:
            for f in self.flat:
                f.div_(self.world)
        self._pending = [len(m) for m in self.members]
        self._done = [False] * self.n_buckets

    def write_back(self) -> None:
        """copy the reduced gradients into ``p.grad`` (for optimisers that read ``.grad``; FusedAdam does not need it)"""
        scale = 1.0 if (self.average or self.world == 1) else 1.0 / self.world
        for p in self.params:
            g = self.views[p]
            if p.grad is None:
                p.grad = torch.empty_like(p)
            p.grad.copy_(g)
            if scale != 1.0:
                p.grad.mul_(scale)

    def describe(self) -> dict:
        return {"buckets": [int(f.numel() * 4) for f in self.flat] if self._built else None, "world": self.world,
                "launched_from_hooks": self.launched_early, "scale": "1/world folded into the Adam kernel" if not self.average else "in place"}

    def remove(self) -> None:
        for h in self._hooks:
            h.remove()


def allreduce_gradients(params, world_size: int | None = None, group=None) -> None:
    """average ``.grad`` of `params` over the process group in place (one flat bucket).  The layout is fixed by `params`
    alone: a parameter whose ``.grad`` is None on this rank contributes zeros and receives the average, so ranks whose
    batches exercise different parameters still reduce identical buffers."""
    if not dist.is_available() or not dist.is_initialized():
        return
    world = world_size or dist.get_world_size(group)
    if world == 1:
        return
    ps = [p for p in params if p.numel() > 0]
    if not ps:
        return
    red = GradientReducer(ps, n_buckets=1, group=group, average=True)
    try:
        red._build()
        red.finish()
        red.write_back()
    finally:
        red.remove()


def shard_indices(n_items: int, rank: int, world_size: int) -> range:
    """DistributedSampler-style contiguous-stride shard of a dataset of n_items structures"""
    return range(rank, n_items, world_size)
