"""Shared test utilities: synthetic molecular / periodic inputs and oracle access."""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
SHIMS = os.path.join(ROOT, "oracle", "shims")
if SHIMS not in sys.path:
    sys.path.insert(0, SHIMS)


def random_molecules(n_mol, n_lo, n_hi, n_species, seed, min_dist=0.95, probs=None):
    """Rejection-sampled blobs: atoms in a sphere of radius 0.95*n^(1/3)*1.1 A with a minimum pair distance."""
    rng = np.random.default_rng(seed)
    pos, z, batch = [], [], []
    for m in range(n_mol):
        n = int(rng.integers(n_lo, n_hi + 1))
        rad = 0.95 * n ** (1.0 / 3.0) * 1.1
        pts = []
        while len(pts) < n:
            p = rng.uniform(-rad, rad, size=3)
            if np.linalg.norm(p) > rad:
                continue
            if all(np.linalg.norm(p - q) >= min_dist for q in pts):
                pts.append(p)
        pos.append(np.array(pts))
        z.append(rng.choice(n_species, size=n, p=probs))
        batch.append(np.full(n, m))
    return (torch.tensor(np.concatenate(pos), dtype=torch.float32), torch.tensor(np.concatenate(z), dtype=torch.long),
            torch.tensor(np.concatenate(batch), dtype=torch.long))


def radius_graph_n2(pos, batch, cutoff):
    """Non-PBC radius graph per molecule: d^2 <= rc^2, i != j, row-major (sender, receiver) order
    (reference neighbourhood_list_torch.py:33-50)."""
    d2 = ((pos.unsqueeze(0) - pos.unsqueeze(1)) ** 2).sum(-1)
    same = batch.unsqueeze(0) == batch.unsqueeze(1)
    mask = (d2 <= cutoff**2) & same
    mask.fill_diagonal_(False)
    return mask.nonzero().T.contiguous()


def make_batch(pos, z, batch, n_species, cutoff, device="cpu", dtype=torch.float32):
    ei = radius_graph_n2(pos, batch, cutoff)
    return {
        "coordinates": pos.to(device=device, dtype=dtype),
        "edge_index": ei.to(device),
        "node_attrs": torch.nn.functional.one_hot(z, n_species).to(device=device, dtype=dtype),
        "batch": batch.to(device),
        "num_graphs": torch.tensor(int(batch.max()) + 1),
    }


def to_device(data, device, dtype=None):
    out = {}
    for k, v in data.items():
        if torch.is_tensor(v):
            if v.is_floating_point() and dtype is not None:
                v = v.to(dtype)
            v = v.to(device) if k != "num_graphs" else v
        out[k] = v
    return out
