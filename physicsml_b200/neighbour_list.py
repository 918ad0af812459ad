"""Drop-in for the reference's neighbour-list entry point, running the GPU cell-list kernels (K1).

``construct_edge_indices_and_attrs`` keeps the reference signature and return convention
(``src/physicsml/lightning/graph_datasets/neighbourhood_list_torch.py:53-61``): ``edge_index [2, E]`` int64 with
``edge_index[0]`` = sender sorted ascending, edge attributes (``None`` unless bond features were supplied, in which case
the reference's merge branch ``:98-156`` is reproduced) and ``cell_shift_vector [E, 3]`` in the position dtype, such that
``r = pos[receiver] - pos[sender] + shift @ cell`` (models/utils.py:77-91).

``radius_graph`` additionally returns the :class:`physicsml_b200.ops.EdgePlan` (int32 edge copies, CSR row pointer and
receiver-grouped permutation) so the model does not have to re-derive it, and accepts a ``batch`` vector to build the
graph of a whole COLLATED BATCH in one call — open-boundary molecules or periodic structures, the
latter either sharing one cell (the reference's batch contract, graph_dataset.py:197-200,385) or with one cell each.
The cell geometry (inverse, bins, stencil) is derived on the device: the only host synchronisation of a call is the
read-back of the edge count that sizes the outputs.  :func:`collate` assembles such a batch from raw structures.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib, ops
from .ops import _call, _p, _stream


def radius_graph(positions: torch.Tensor, cutoff: float, pbc=None, cell: Optional[torch.Tensor] = None,
                 self_interaction: bool = False, batch: Optional[torch.Tensor] = None, num_graphs: Optional[int] = None,
                 cartesian_shifts: bool = False):
    """-> (edge_index int64 [2,E], shift f32 [E,3], EdgePlan)

    cell: ``[3,3]`` (shared by all structures of the batch) or ``[G,3,3]`` (one per structure).  With per-structure
    cells pass ``cartesian_shifts=True``: the returned shift rows are then ``S @ cell_g`` in length units, to be used
    with ``data["cell"] = torch.eye(3)`` — the reference's model contract knows only one cell per batch."""
    if not positions.is_cuda:
        raise RuntimeError("physicsml_b200 neighbour list runs on CUDA tensors only (no CPU fallback)")
    pos = positions.detach().to(torch.float32).contiguous()
    N = pos.shape[0]
    dev = pos.device
    periodic = pbc is not None and all(bool(x) for x in pbc) and cell is not None
    if pbc is not None and any(bool(x) for x in pbc) and not periodic:
        raise NotImplementedError("mixed periodic / open boundaries are not supported (the reference's linked cell is not either)")
    i32 = dict(dtype=torch.int32, device=dev)
    deg = torch.empty(N, **i32)
    rowptr = torch.empty(N + 1, **i32)
    fl = None
    if batch is None:
        G, b64 = 1, None
        gptr = torch.tensor([0, N], **i32)
    else:
        b64 = batch.to(torch.int64).contiguous()
        # pass num_graphs to avoid a device read-back here
        G = int(num_graphs) if num_graphs is not None else int(b64.max().item()) + 1
        gptr = torch.zeros(G + 1, **i32)
        gptr[1:] = torch.cumsum(torch.bincount(b64, minlength=G), 0)
    rc2 = float(np.float32(cutoff * cutoff))
    err = torch.zeros(2, **i32)  # [0]: segment longer than the sort supports, [1]: setup flags
    if periodic:
        cells = cell.detach().to(device=dev, dtype=torch.float32).reshape(-1, 3, 3).contiguous()
        if cells.shape[0] not in (1, G):
            raise ValueError(f"cell must be [3,3] or [{G},3,3], got {tuple(cell.shape)}")
        if cells.shape[0] > 1 and not cartesian_shifts:
            raise ValueError("per-structure cells need cartesian_shifts=True (the model contract has one cell per batch)")
        stride = 9 if cells.shape[0] > 1 else 0
        graphs = torch.empty(G * _lib.NL_GRAPH_BYTES, dtype=torch.uint8, device=dev)
        nbins_t = torch.zeros(1, **i32)
        _call("pml_nl_setup", _p(cells), stride, _p(gptr), G, float(cutoff), _p(graphs), _p(nbins_t), _p(err[1:]), _stream())
        max_bins = 64 * G + 4 * N  # bound of what pml_nl_setup hands out
        wpos = torch.empty(N, 3, dtype=torch.float32, device=dev)
        fl = torch.empty(N, 3, **i32)
        bins = torch.empty(N, **i32)
        scratch = torch.zeros(2, max_bins, **i32)  # bincount, cursor
        binptr = torch.empty(max_bins + 1, **i32)
        order = torch.empty(N, **i32)
        _call("pml_nl_pbc_bin", _p(pos), _p(b64), _p(graphs), _p(wpos), _p(fl), _p(bins), _p(scratch[0]), N, _stream())
        _call("pml_scan_i32", _p(scratch[0]), _p(binptr), max_bins, _stream())
        _call("pml_nl_pbc_fill_bins", _p(bins), _p(binptr), _p(scratch[1]), _p(order), N, _stream())
        _call("pml_nl_pbc_count", _p(wpos), _p(b64), _p(bins), _p(binptr), _p(order), _p(graphs), rc2, int(self_interaction),
              _p(deg), N, _stream())
    else:
        _call("pml_nl_open_count", _p(pos), _p(b64), _p(gptr), rc2, int(self_interaction), _p(deg), N, _stream())
    _call("pml_scan_i32", _p(deg), _p(rowptr), N, _stream())
    E = int(rowptr[-1].item())  # the only host sync: the edge count sizes the outputs
    keys = torch.empty(max(E, 1), dtype=torch.int64, device=dev)
    if periodic:
        _call("pml_nl_pbc_fill", _p(wpos), _p(b64), _p(bins), _p(binptr), _p(order), _p(graphs), rc2, int(self_interaction),
              _p(rowptr), _p(keys), N, _stream())
        _call("pml_nl_sort_segments", _p(keys), _p(rowptr), N, _p(err), _stream())
    else:
        _call("pml_nl_open_fill", _p(pos), _p(b64), _p(gptr), rc2, int(self_interaction), _p(rowptr), _p(keys), N, _stream())
    edge_index = torch.empty(2, E, dtype=torch.int64, device=dev)
    shift = torch.zeros(E, 3, dtype=torch.float32, device=dev)
    snd = torch.empty(E, **i32)
    rcv = torch.empty(E, **i32)
    perm = torch.empty(E, **i32)
    _call("pml_nl_emit", _p(keys), _p(rowptr), _p(fl), _p(edge_index), _p(shift) if periodic else None, _p(snd), _p(rcv),
          _p(perm), N, E, _stream())
    if periodic:
        flags = err.tolist()  # after the emit: this sync costs nothing extra, the caller consumes the edges next
        if flags[0] != 0:
            raise RuntimeError("an atom has more than 2048 neighbours; the per-sender sort does not support that")
        if flags[1] & 2:
            raise ValueError("cutoff spans more than 127 periodic images; unsupported")
        if cartesian_shifts:  # S @ cell of the sender's structure (plain index plumbing on the device)
            cg = cells[b64[edge_index[0]]] if cells.shape[0] > 1 else cells[0].expand(E, 3, 3)
            shift = torch.bmm(shift.unsqueeze(1), cg).squeeze(1)
    plan_out = ops.EdgePlan.from_sorted(snd, rcv, rowptr, perm, N)
    plan_out._src = (edge_index.data_ptr(), edge_index._version, tuple(edge_index.shape))
    return edge_index, shift, plan_out


def _coalesce_add(index: torch.Tensor, values, num_nodes: int):
    """torch_geometric.utils.coalesce(edge_index, [values...], num_nodes) with its default reduce="sum": edges sorted
    by (row, col), duplicates merged by adding their value rows"""
    key = index[0] * num_nodes + index[1]
    uniq, inv = torch.unique(key, sorted=True, return_inverse=True)
    outs = []
    for v in values:
        o = torch.zeros((uniq.numel(),) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device)
        o.index_add_(0, inv, v)
        outs.append(o)
    return torch.stack([uniq // num_nodes, uniq % num_nodes]), outs


def _merge_bond_edges(edge_index, shift, initial_edge_indices, initial_edge_attrs, num_nodes: int):
    """The reference's bond-feature merge (neighbourhood_list_torch.py:98-156): the supplied edges ([B, 2] rows) are made
    undirected (PyG ``to_undirected``: both directions, duplicates summed), the radius-graph edges get all-zero attribute
    rows and the supplied edges zero shifts, everything is coalesced by (sender, receiver) with "add", and the supplied
    edges must be a subset of the radius graph (the reference asserts the edge count does not change).  Index plumbing
    with stock device ops — this branch runs once per structure in data preparation, not in the message-passing path."""
    dev = edge_index.device
    bi = initial_edge_indices.to(dev).T.contiguous()
    ba = initial_edge_attrs.to(dev)
    bi, (ba,) = _coalesce_add(torch.cat([bi, bi.flip(0)], dim=1), [torch.cat([ba, ba], dim=0)], num_nodes)
    all_index = torch.cat([edge_index, bi], dim=1)
    all_attrs = torch.cat([torch.zeros(edge_index.shape[1], ba.shape[1], dtype=ba.dtype, device=dev), ba], dim=0)
    all_shift = torch.cat([shift, torch.zeros(bi.shape[1], 3, dtype=shift.dtype, device=dev)], dim=0)
    out_index, (out_attrs, out_shift) = _coalesce_add(all_index, [all_attrs, all_shift], num_nodes)
    if out_index.shape != edge_index.shape:
        raise ValueError("the supplied (bond) edges are not a subset of the radius graph, or an atom pair is connected "
                         "through several periodic images (the reference asserts the same, neighbourhood_list_torch.py:155)")
    return out_index, out_attrs, out_shift


def construct_edge_indices_and_attrs(positions: torch.Tensor, cutoff: float, initial_edge_indices, initial_edge_attrs, pbc,
                                     cell, self_interaction: bool):
    """reference signature (neighbourhood_list_torch.py:53-61) -> (edge_index, edge_attrs or None, cell_shift_vector).
    ``initial_edge_indices`` is [B, 2] as in the reference (one row per supplied edge)."""
    edge_index, shift, _ = radius_graph(positions, cutoff, pbc=pbc, cell=cell, self_interaction=self_interaction)
    shift = shift.to(positions.dtype)
    if initial_edge_indices is not None and initial_edge_attrs is not None:
        if initial_edge_attrs.dim() > 1:
            edge_index, attrs, shift = _merge_bond_edges(edge_index, shift, initial_edge_indices, initial_edge_attrs,
                                                         positions.shape[0])
            return edge_index, attrs, shift
        return edge_index, initial_edge_attrs, shift  # neighbourhood_list_torch.py:157-164
    return edge_index, None, shift


def collate(structures: Sequence[dict], cutoff: float, device, num_species: Optional[int] = None,
            self_interaction: bool = False) -> dict:
    """Batch assembly on the device (replaces the per-sample CPU neighbour list in ``GraphDataset.get`` +
    PyG collate, graph_dataset.py:247-386, datamodule.py:72-110): raw structures in, the reference's batch dict out, with
    ONE neighbour-list call for the whole batch.

    structures: dicts with ``coordinates [n,3]``, and either ``node_attrs [n,Z]`` or ``species [n]`` (+ ``num_species``);
    optional ``cell [3,3]`` (all periodic or none; cells may differ per structure), ``total_atomic_energy``.
    -> dict(coordinates, node_attrs, edge_index, batch, num_graphs, ptr [, cell, cell_shift_vector, total_atomic_energy],
            _edge_plan)"""
    dev = torch.device(device)
    ns = [int(s["coordinates"].shape[0]) for s in structures]
    G = len(structures)
    pos = torch.cat([torch.as_tensor(s["coordinates"], dtype=torch.float32) for s in structures]).to(dev, non_blocking=True)
    if "node_attrs" in structures[0]:
        attrs = torch.cat([torch.as_tensor(s["node_attrs"], dtype=torch.float32) for s in structures]).to(dev, non_blocking=True)
    else:
        if num_species is None:
            raise ValueError("pass num_species with species indices")
        z = torch.cat([torch.as_tensor(s["species"], dtype=torch.long) for s in structures]).to(dev, non_blocking=True)
        attrs = torch.nn.functional.one_hot(z, num_species).to(torch.float32)
    batch = torch.repeat_interleave(torch.arange(G), torch.tensor(ns)).to(dev, non_blocking=True)
    ptr = torch.zeros(G + 1, dtype=torch.long)
    ptr[1:] = torch.cumsum(torch.tensor(ns), 0)
    data = {"coordinates": pos, "node_attrs": attrs, "batch": batch, "num_graphs": torch.tensor(G), "ptr": ptr.to(dev)}
    periodic = "cell" in structures[0] and structures[0]["cell"] is not None
    if periodic:
        cells = torch.stack([torch.as_tensor(s["cell"], dtype=torch.float32).reshape(3, 3) for s in structures])
        shared = bool((cells == cells[0]).all())
        if shared:
            ei, sh, plan = radius_graph(pos, cutoff, pbc=(True, True, True), cell=cells[0].to(dev), batch=batch, num_graphs=G,
                                        self_interaction=self_interaction)
            data["cell"] = cells[0].to(dev)
        else:
            ei, sh, plan = radius_graph(pos, cutoff, pbc=(True, True, True), cell=cells.to(dev), batch=batch, num_graphs=G,
                                        self_interaction=self_interaction, cartesian_shifts=True)
            data["cell"] = torch.eye(3, device=dev)  # shifts are already in length units
        data["cell_shift_vector"] = sh
    else:
        ei, sh, plan = radius_graph(pos, cutoff, batch=batch, num_graphs=G, self_interaction=self_interaction)
    data["edge_index"] = ei
    data["_edge_plan"] = plan
    if "total_atomic_energy" in structures[0]:
        data["total_atomic_energy"] = torch.tensor([float(s["total_atomic_energy"]) for s in structures]).to(dev)
    return data
