"""Synthetic workloads of the BASELINE.json configs: deterministic inputs, model hyper-parameters and
the training step (forward, forces with create_graph, MSE(E)+MSE(F), backward, Adam-amsgrad) shared by bench.py,
the tests and the CPU baseline.  Pure host-side data construction (numpy / torch CPU); no kernels here.
"""
from __future__ import annotations

import numpy as np
import torch

CONFIGS = {
    # configs[0]: MACE small inference, 64 QM9-sized molecules
    "mace_small_qm9": dict(
        kind="inference", n_mol=64, atoms=(12, 24), Z=5, seed=101, probs=[0.50, 0.35, 0.06, 0.08, 0.01], pbc=False,
        model=dict(cut_off=5.0, num_bessel=8, num_polynomial_cutoff=5, max_ell=2, num_interactions=2, num_node_feats=5,
                   num_edge_feats=0, hidden_irreps="32x0e + 32x1o", correlation=3), mlp="16x0e"),
    # configs[1]: MACE medium training, 32 SPICE-like molecules per GPU  (the metric's workload)
    "mace_medium_spice_train": dict(
        kind="training", n_mol=32, atoms=(30, 50), Z=10, seed=202, probs=None, pbc=False,
        model=dict(cut_off=5.0, num_bessel=8, num_polynomial_cutoff=5, max_ell=3, num_interactions=2, num_node_feats=10,
                   num_edge_feats=0, hidden_irreps="128x0e + 128x1o", correlation=3), mlp="16x0e"),
}


def water_box(n_mol: int = 8000, box: float = 62.14, seed: int = 404):
    """configs[3]: n_mol water molecules on a jittered cubic lattice in a periodic box (rho = 0.1 atoms/A^3 at the
    defaults), O-H 0.96 A, H-O-H 104.5 deg, random orientation.  -> (pos [3 n_mol, 3], species idx (O=3, H=0), cell)"""
    rng = np.random.default_rng(seed)
    m = int(np.ceil(n_mol ** (1.0 / 3.0)))
    a = box / m
    idx = np.stack(np.meshgrid(np.arange(m), np.arange(m), np.arange(m), indexing="ij"), -1).reshape(-1, 3)[:n_mol]
    centres = (idx + 0.5) * a + rng.uniform(-0.25, 0.25, size=(n_mol, 3)) * (a - 2.6)
    # random rotations (QR of gaussian matrices)
    q, r = np.linalg.qr(rng.normal(size=(n_mol, 3, 3)))
    q = q * np.sign(np.diagonal(r, axis1=1, axis2=2))[:, None, :]
    th = np.deg2rad(104.5) / 2
    h1 = 0.96 * np.array([np.sin(th), np.cos(th), 0.0])
    h2 = 0.96 * np.array([-np.sin(th), np.cos(th), 0.0])
    pos = np.stack([centres, centres + q @ h1, centres + q @ h2], axis=1).reshape(-1, 3)
    z = np.tile(np.array([3, 0, 0]), n_mol)
    return torch.tensor(pos, dtype=torch.float32), torch.tensor(z, dtype=torch.long), torch.eye(3) * box


def random_molecules(n_mol, n_lo, n_hi, n_species, seed, min_dist=0.95, probs=None):
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
    """Reference semantics of the non-PBC radius graph (neighbourhood_list_torch.py:33-50), per molecule."""
    out = []
    for m in range(int(batch.max()) + 1):
        idx = (batch == m).nonzero().flatten()
        p = pos[idx]
        d2 = ((p.unsqueeze(0) - p.unsqueeze(1)) ** 2).sum(-1)
        mask = d2 <= cutoff**2
        mask.fill_diagonal_(False)
        e = mask.nonzero()
        out.append(idx[e].T)
    return torch.cat(out, dim=1).contiguous()


def make_batch(name: str, rank: int = 0, n_mol: int | None = None):
    """-> (batch dict on CPU, targets dict, model kwargs incl. measured avg_num_neighbours, mlp_irreps)"""
    cfg = CONFIGS[name]
    n_mol = n_mol or cfg["n_mol"]
    pos, z, batch = random_molecules(n_mol, cfg["atoms"][0], cfg["atoms"][1], cfg["Z"], cfg["seed"] + 1000 * rank, probs=cfg["probs"])
    ei = radius_graph_n2(pos, batch, cfg["model"]["cut_off"])
    data = {
        "coordinates": pos,
        "edge_index": ei,
        "node_attrs": torch.nn.functional.one_hot(z, cfg["Z"]).float(),
        "batch": batch,
        "num_graphs": torch.tensor(n_mol),
    }
    g = torch.Generator().manual_seed(cfg["seed"] + 7 + rank)
    targets = {"energy": torch.randn(n_mol, 1, generator=g), "forces": 0.1 * torch.randn(pos.shape[0], 3, generator=g)}
    # avg_num_neighbours is a model hyper-parameter: fixed from the rank-0 reference batch so all ranks share weights
    model_kw = dict(cfg["model"])
    return data, targets, model_kw, cfg["mlp"]


def mean_degree(name: str) -> float:
    data, _, _, _ = make_batch(name, 0)
    return data["edge_index"].shape[1] / data["coordinates"].shape[0]


def training_loss(model, data: dict, targets: dict) -> torch.Tensor:
    """forward -> forces (create_graph) -> MSE(E) + MSE(F)   (reference lightning/module.py:65-95, mace_module.py:175-185)"""
    d = dict(data)
    pos = d["coordinates"].detach().requires_grad_(True)
    d["coordinates"] = pos
    e = model(d)["y_graph_scalars"]
    from . import ops

    with ops.input_grads_only():  # the force pass: gradients w.r.t. the coordinates only
        (g,) = torch.autograd.grad([e], [pos], grad_outputs=[torch.ones_like(e)], retain_graph=True, create_graph=True)
    return torch.nn.functional.mse_loss(e, targets["energy"]) + torch.nn.functional.mse_loss(-g, targets["forces"])


def make_optimizer(model, capturable: bool = False):
    """reference default: Adam(lr=1e-2, amsgrad=True, weight_decay=5e-7) (mace/supervised/default_configs.py:25-33).
    CUDA parameters get the fused single-kernel implementation of the same update rule (physicsml_b200.optim.FusedAdam,
    always graph-capturable); CPU parameters (the oracle / reference arm) get torch.optim.Adam."""
    params = [p for p in model.parameters() if p.requires_grad and p.numel() > 0]
    if params and params[0].is_cuda:
        from .optim import FusedAdam

        return FusedAdam(params, lr=1e-2, weight_decay=5e-7, amsgrad=True)
    return torch.optim.Adam(params, lr=1e-2, amsgrad=True, weight_decay=5e-7, capturable=capturable)
