"""CPU restatement of the reference's radius-graph neighbour list (edge SET semantics).

TEST INFRASTRUCTURE ONLY (oracle): only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.

Restates, in plain fp32 torch on the CPU (reference paths under src/physicsml/lightning/graph_datasets):
  * open boundaries  — compute_neighborlist_n2_no_cell, neighbourhood_list_torch.py:33-50: dense N x N, d^2 <= rc^2, i != j,
    row-major (sender, receiver) order, zero shifts;
  * full PBC         — construct_edge_indices_and_attrs, neighbourhood_list_torch.py:63-88 (wrap by floor(pos @ inv(cell)),
    shift fix-up S += floor_i - floor_j) around compute_neighborlist (torch_nl_vendored/neighbor_list.py:108-164):
    images within num_repeats = ceil(rc * |rows of inv(cell)^T|) (utils.py:41-57), strict d^2 < rc^2 with
    d = pos_i - pos_j - S.cell (neighbor_list.py:38-44), same-image self pairs dropped (linked_cell.py:273-275).
The linked-cell machinery of the reference only accelerates the search; the set it returns is the brute-force set
enumerated here (checked against the reference itself in tests/golden/make_golden.py).

Output: edges sorted by (sender, receiver, shift) — the reference guarantees only the sender order
(linked_cell.py:277-279), so comparisons are made after sorting, as BASELINE.json's north_star specifies.
"""
from __future__ import annotations

import itertools

import torch


def _sort_edges(i, j, S):
    if i.numel() == 0:
        return torch.stack([i, j]), S
    Sl = S.to(torch.int64)
    lo = Sl.min()
    key = ((i * (j.max() + 1) + j) * 512 + (Sl[:, 0] - lo)) * 512 * 512 + (Sl[:, 1] - lo) * 512 + (Sl[:, 2] - lo)
    order = torch.argsort(key, stable=True)
    return torch.stack([i[order], j[order]]), S[order]


def ref_radius_graph(pos: torch.Tensor, cutoff: float, pbc=None, cell=None, self_interaction: bool = False):
    """-> (edge_index int64 [2,E] sorted by (sender, receiver, shift), cell_shift_vector [E,3] in pos.dtype)"""
    pos = pos.detach().cpu()
    n = pos.shape[0]
    if pbc is not None and all(bool(x) for x in pbc) and cell is not None:
        cell = cell.detach().cpu().to(pos.dtype).reshape(-1, 3)[:3]
        inv = torch.inverse(cell)
        coefs = torch.floor(pos @ inv)
        w = pos - coefs @ cell
        reps = torch.ceil(cutoff * torch.linalg.inv(cell).transpose(1, 0).norm(2, dim=-1)).to(torch.long).tolist()
        out_i, out_j, out_s = [], [], []
        rc2 = cutoff * cutoff
        ar = torch.arange(n)
        for s in itertools.product(*[range(-r, r + 1) for r in reps]):
            sv = torch.tensor(s, dtype=pos.dtype)
            shift_cart = sv @ cell
            d2 = (w[:, None, :] - w[None, :, :] - shift_cart).square().sum(-1)
            mask = d2 < rc2
            if s == (0, 0, 0) and not self_interaction:
                mask[ar, ar] = False
            ij = mask.nonzero()
            out_i.append(ij[:, 0])
            out_j.append(ij[:, 1])
            out_s.append(sv.expand(ij.shape[0], 3))
        i, j, S = torch.cat(out_i), torch.cat(out_j), torch.cat(out_s)
        S = S + (coefs[i] - coefs[j])
        return _sort_edges(i, j, S)
    d2 = ((pos.unsqueeze(0) - pos.unsqueeze(1)) ** 2).sum(-1)
    ij = (d2 <= cutoff**2).nonzero()
    if not self_interaction:
        ij = ij[ij[:, 0] != ij[:, 1]]
    return ij.T.contiguous(), torch.zeros(ij.shape[0], 3, dtype=pos.dtype)


def ref_radius_graph_batched(pos, batch, cutoff, self_interaction: bool = False):
    """The reference builds one graph per sample and PyG collates them (graph_dataset.py:314-322 + Batch collate):
    edge indices are offset by the node count of the preceding samples."""
    eis = []
    for g in range(int(batch.max()) + 1):
        idx = (batch == g).nonzero().flatten()
        ei, _ = ref_radius_graph(pos[idx], cutoff, None, None, self_interaction)
        eis.append(idx[ei])
    ei = torch.cat(eis, dim=1)
    return ei, torch.zeros(ei.shape[1], 3, dtype=pos.dtype)


def sorted_edge_set(edge_index, shift):
    """canonical order for set comparison"""
    return _sort_edges(edge_index[0].cpu(), edge_index[1].cpu(), shift.cpu())
