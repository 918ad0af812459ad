"""Drop-in for the reference's neighbour-list entry point, running the GPU cell-list kernels (K1).

``construct_edge_indices_and_attrs`` keeps the reference signature and return convention
(``src/physicsml/lightning/graph_datasets/neighbourhood_list_torch.py:53-61``): ``edge_index [2, E]`` int64 with
``edge_index[0]`` = sender sorted ascending, ``None`` edge attributes and ``cell_shift_vector [E, 3]`` in the position
dtype, such that ``r = pos[receiver] - pos[sender] + shift @ cell`` (models/utils.py:77-91).  The bond-feature merge
branch (neighbourhood_list_torch.py:98-156) is out of scope.

``radius_graph`` additionally returns the :class:`physicsml_b200.ops.EdgePlan` (int32 edge copies, CSR row pointer and
receiver-grouped permutation) so the model does not have to re-derive it, and accepts a ``batch`` vector to build the
graph of a whole collated batch of open-boundary molecules in one call.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import numpy as np
import torch

from . import _lib, ops
from .ops import _call, _p, _stream


def _cell_plan(cell: torch.Tensor, cutoff: float, n_atoms: int) -> _lib.NLCell:
    c = cell.detach().to("cpu", torch.float64).reshape(-1, 3)[:3].numpy()
    inv = np.linalg.inv(c)
    # perpendicular height of the cell along axis k = 1 / |column k of inv|  (utils.py:41-57 uses the same quantity)
    h = 1.0 / np.linalg.norm(inv, axis=0)
    st = _lib.NLCell()
    max_bins = max(64, 4 * n_atoms)
    nb = [max(1, min(128, int(math.floor(hk / cutoff)))) for hk in h]
    while nb[0] * nb[1] * nb[2] > max_bins:
        k = int(np.argmax(nb))
        nb[k] = max(1, nb[k] // 2)
    for k in range(3):
        st.nb[k] = nb[k]
        st.R[k] = int(math.ceil(cutoff * nb[k] / h[k] - 1e-12))
        if st.R[k] > 127:
            raise ValueError("cutoff spans more than 127 periodic images; unsupported")
    for k in range(9):
        st.inv[k] = float(inv.reshape(-1)[k])
        st.cell[k] = float(np.float32(c.reshape(-1)[k]))
    st.rc2 = float(np.float32(cutoff * cutoff))
    return st


def radius_graph(positions: torch.Tensor, cutoff: float, pbc=None, cell: Optional[torch.Tensor] = None,
                 self_interaction: bool = False, batch: Optional[torch.Tensor] = None, num_graphs: Optional[int] = None):
    """-> (edge_index int64 [2,E], shift f32 [E,3], EdgePlan)"""
    if not positions.is_cuda:
        raise RuntimeError("physicsml_b200 neighbour list runs on CUDA tensors only (no CPU fallback)")
    pos = positions.detach().to(torch.float32).contiguous()
    N = pos.shape[0]
    dev = pos.device
    periodic = pbc is not None and all(bool(x) for x in pbc) and cell is not None
    i32 = dict(dtype=torch.int32, device=dev)
    deg = torch.empty(N, **i32)
    rowptr = torch.empty(N + 1, **i32)
    fl = None
    if periodic:
        if batch is not None and int(batch.max()) > 0:
            raise NotImplementedError("periodic neighbour lists are built one structure at a time (one cell per batch)")
        plan = _cell_plan(cell, cutoff, N)
        nbins = plan.nb[0] * plan.nb[1] * plan.nb[2]
        wpos = torch.empty(N, 3, dtype=torch.float32, device=dev)
        fl = torch.empty(N, 3, **i32)
        bins = torch.empty(N, **i32)
        bincount = torch.zeros(nbins, **i32)
        binptr = torch.empty(nbins + 1, **i32)
        cursor = torch.zeros(nbins, **i32)
        order = torch.empty(N, **i32)
        _call("pml_nl_pbc_bin", _p(pos), C.byref(plan), _p(wpos), _p(fl), _p(bins), _p(bincount), N, _stream())
        _call("pml_scan_i32", _p(bincount), _p(binptr), nbins, _stream())
        _call("pml_nl_pbc_fill_bins", _p(bins), _p(binptr), _p(cursor), _p(order), N, _stream())
        _call("pml_nl_pbc_count", _p(wpos), _p(bins), _p(binptr), _p(order), C.byref(plan), int(self_interaction), _p(deg), N,
              _stream())
    else:
        if batch is None:
            gptr = torch.tensor([0, N], **i32)
            b64 = None
        else:
            b64 = batch.to(torch.int64).contiguous()
            G = int(num_graphs) if num_graphs is not None else int(b64.max()) + 1
            cnt = torch.bincount(b64, minlength=G).to(torch.int32)
            gptr = torch.zeros(G + 1, **i32)
            gptr[1:] = torch.cumsum(cnt, 0)
        rc2 = float(np.float32(cutoff**2))
        _call("pml_nl_open_count", _p(pos), _p(b64), _p(gptr), rc2, int(self_interaction), _p(deg), N, _stream())
    _call("pml_scan_i32", _p(deg), _p(rowptr), N, _stream())
    E = int(rowptr[-1].item())  # the only host sync: the edge count sizes the outputs
    keys = torch.empty(max(E, 1), dtype=torch.int64, device=dev)
    if periodic:
        _call("pml_nl_pbc_fill", _p(wpos), _p(bins), _p(binptr), _p(order), C.byref(plan), int(self_interaction), _p(rowptr),
              _p(keys), N, _stream())
        err = torch.zeros(1, **i32)
        _call("pml_nl_sort_segments", _p(keys), _p(rowptr), N, _p(err), _stream())
    else:
        _call("pml_nl_open_fill", _p(pos), _p(b64), _p(gptr), rc2, int(self_interaction), _p(rowptr), _p(keys), N, _stream())
        err = None
    edge_index = torch.empty(2, E, dtype=torch.int64, device=dev)
    shift = torch.zeros(E, 3, dtype=torch.float32, device=dev)
    snd = torch.empty(E, **i32)
    rcv = torch.empty(E, **i32)
    perm = torch.empty(E, **i32)
    _call("pml_nl_emit", _p(keys), _p(rowptr), _p(fl), _p(edge_index), _p(shift) if periodic else None, _p(snd), _p(rcv),
          _p(perm), N, E, _stream())
    if err is not None and int(err.item()) != 0:
        raise RuntimeError("an atom has more than 2048 neighbours; the per-sender sort does not support that")
    plan_out = ops.EdgePlan.from_sorted(snd, rcv, rowptr, perm, N)
    plan_out._src = (edge_index.data_ptr(), edge_index._version, tuple(edge_index.shape))
    return edge_index, shift, plan_out


def construct_edge_indices_and_attrs(positions: torch.Tensor, cutoff: float, initial_edge_indices, initial_edge_attrs, pbc,
                                     cell, self_interaction: bool):
    """reference signature (neighbourhood_list_torch.py:53-61) -> (edge_index, edge_attrs=None, cell_shift_vector)"""
    if initial_edge_indices is not None and initial_edge_attrs is not None and initial_edge_attrs.dim() > 1:
        raise NotImplementedError("merging bond-feature edges is out of scope of the B200 hot path (SURVEY.md §2 row 1)")
    edge_index, shift, _ = radius_graph(positions, cutoff, pbc=pbc, cell=cell, self_interaction=self_interaction)
    attrs = initial_edge_attrs if (initial_edge_indices is not None and initial_edge_attrs is not None) else None
    return edge_index, attrs, shift.to(positions.dtype)
