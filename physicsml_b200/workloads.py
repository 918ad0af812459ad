"""Synthetic workloads of the five BASELINE.json configs: deterministic inputs, model hyper-parameters
and the evaluation / training step (forward, forces with create_graph, MSE(E)+MSE(F), backward, Adam-amsgrad) shared
by bench.py, the tests and the CPU baseline.  Pure host-side data construction (numpy / torch CPU); no kernels here.

    mace_small_qm9           configs[0]  MACE small   inference   64 QM9-sized molecules, open boundary
    mace_medium_spice_train  configs[1]  MACE medium  training    32 SPICE-like molecules per GPU     (the metric's workload)
    nequip_organic           configs[2]  NequIP 4x64  inference   64 organic molecules of ~30 atoms
    water_box_infer          configs[3]  MACE medium  inference   24 000-atom periodic water box, r_cut 6 A
    mace_large_bulk_train    configs[4]  MACE large   training    8 periodic rocksalt-like cells of 216 atoms per GPU
"""
from __future__ import annotations

import math

import numpy as np
import torch

_MACE_MEDIUM = dict(num_bessel=8, num_polynomial_cutoff=5, max_ell=3, num_interactions=2, num_node_feats=10,
                    num_edge_feats=0, hidden_irreps="128x0e + 128x1o", correlation=3)

CONFIGS = {
    "mace_small_qm9": dict(
        family="mace", kind="inference", n_mol=64, atoms=(12, 24), Z=5, seed=101, probs=[0.50, 0.35, 0.06, 0.08, 0.01], pbc=False,
        model=dict(cut_off=5.0, num_bessel=8, num_polynomial_cutoff=5, max_ell=2, num_interactions=2, num_node_feats=5,
                   num_edge_feats=0, hidden_irreps="32x0e + 32x1o", correlation=3), mlp="16x0e"),
    "mace_medium_spice_train": dict(
        family="mace", kind="training", n_mol=32, atoms=(30, 50), Z=10, seed=202, probs=None, pbc=False,
        model=dict(cut_off=5.0, **_MACE_MEDIUM), mlp="16x0e"),
    "nequip_organic": dict(
        family="nequip", kind="inference", n_mol=64, atoms=(24, 36), Z=5, seed=303, probs=[0.50, 0.35, 0.06, 0.08, 0.01], pbc=False,
        model=dict(cut_off=5.0, num_layers=4, max_ell=2, parity=True, num_features=64, num_bessel=8, bessel_basis_trainable=True,
                   num_polynomial_cutoff=6, self_connection=True, resnet=True, num_node_feats=5, num_edge_feats=0), mlp="16x0e"),
    "water_box_infer": dict(
        family="mace", kind="inference", n_mol=8000, Z=10, seed=404, pbc=True, box=62.14,
        model=dict(cut_off=6.0, **_MACE_MEDIUM), mlp="16x0e", avg_num_neighbours=90.0),
    "mace_large_bulk_train": dict(
        family="mace", kind="training", n_mol=8, Z=4, seed=505, pbc=True, reps=3, lattice=4.5, sigma=0.05,
        model=dict(cut_off=5.0, num_bessel=8, num_polynomial_cutoff=5, max_ell=3, num_interactions=2, num_node_feats=4,
                   num_edge_feats=0, hidden_irreps="256x0e + 256x1o + 256x2e", correlation=3), mlp="16x0e"),
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


def rocksalt_cells(n_cells: int, reps: int = 3, lattice: float = 4.5, sigma: float = 0.05, n_species: int = 4, seed: int = 505,
                   avoid_cutoff: float | None = None):
    """configs[4]: n_cells rocksalt supercells (reps^3 conventional cells = 8 reps^3 atoms each; 216 atoms in a 13.5 A
    cube at the defaults, rho = 0.088 atoms/A^3), Gaussian displacements, two cation and two anion species placed at
    random on their sublattices.  All structures share ONE cell (the reference's batch contract).
    ``avoid_cutoff``: displacements are redrawn until no pair distance lies within 1e-4 A of that cut-off, so that edge-set
    equality between implementations is not at the mercy of rounding.
    -> (pos [n_cells * n, 3], species [n_cells * n], batch, cell [3, 3])"""
    rng = np.random.default_rng(seed)
    base = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5],  # cations (fcc)
                     [0.5, 0, 0], [0, 0.5, 0], [0, 0, 0.5], [0.5, 0.5, 0.5]])  # anions
    kind = np.array([0, 0, 0, 0, 1, 1, 1, 1])
    cells = np.stack(np.meshgrid(np.arange(reps), np.arange(reps), np.arange(reps), indexing="ij"), -1).reshape(-1, 3)
    frac = (cells[:, None, :] + base[None, :, :]).reshape(-1, 3)
    sub = np.tile(kind, len(cells))
    half = max(1, n_species // 2)
    pos, z, batch = [], [], []
    box = reps * lattice
    imgs = np.stack(np.meshgrid(*[np.arange(-1, 2)] * 3, indexing="ij"), -1).reshape(-1, 3) * box
    for c in range(n_cells):
        for _attempt in range(500):
            p = frac * lattice + rng.normal(scale=sigma, size=frac.shape)
            if avoid_cutoff is None:
                break
            d = np.sqrt(((p[None, :, None, :] + imgs[None, None, :, :] - p[:, None, None, :]) ** 2).sum(-1))
            if np.abs(d - avoid_cutoff).min() > 1e-4:
                break
        else:
            raise RuntimeError("could not place a cell without a pair at the cut-off")
        pos.append(p)
        z.append(np.where(sub == 0, rng.integers(0, half, size=len(sub)), half + rng.integers(0, n_species - half, size=len(sub))))
        batch.append(np.full(len(sub), c))
    cell = torch.eye(3) * (reps * lattice)
    return (torch.tensor(np.concatenate(pos), dtype=torch.float32), torch.tensor(np.concatenate(z), dtype=torch.long),
            torch.tensor(np.concatenate(batch), dtype=torch.long), cell)


def random_molecules(n_mol, n_lo, n_hi, n_species, seed, min_dist=0.95, probs=None, sizes=None):
    """sizes: atom counts to use instead of drawing them (the size draw is still consumed, so that a batch built with its
    own sizes is reproduced bit for bit)"""
    rng = np.random.default_rng(seed)
    pos, z, batch = [], [], []
    for m in range(n_mol):
        n = int(rng.integers(n_lo, n_hi + 1))
        if sizes is not None:
            n = int(sizes[m])
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


def radius_graph_pbc_small(pos, batch, cell, cutoff):
    """Brute-force periodic radius graph for SMALL structures sharing one cell (host-side data construction for the
    configs[4] batches and the CPU baseline; the product's neighbour list is physicsml_b200.neighbour_list).  Same
    conventions as the reference (neighbourhood_list_torch.py:63-96): positions wrapped into the cell, strict
    d^2 < rc^2, r = pos[j] - pos[i] + S @ cell for edge (i, j, S); edges sorted by (sender, receiver, S).
    -> (edge_index [2, E] int64, shift [E, 3] float32)"""
    c = cell.double().numpy()
    inv = np.linalg.inv(c)
    h = 1.0 / np.linalg.norm(inv, axis=0)
    R = [int(math.ceil(cutoff / hk)) for hk in h]
    imgs = np.stack(np.meshgrid(*[np.arange(-r, r + 1) for r in R], indexing="ij"), -1).reshape(-1, 3)
    p64 = pos.double().numpy()
    eis, shs = [], []
    for m in range(int(batch.max()) + 1):
        idx = (batch == m).nonzero().flatten().numpy()
        p = p64[idx]
        coef = np.floor(p @ inv)
        w = p - coef @ c  # wrapped positions
        i_all, j_all, s_all = [], [], []
        for S in imgs:
            d = w[None, :, :] + (S @ c)[None, None, :] - w[:, None, :]  # [i, j]: w_j + S c - w_i
            d2 = (d.astype(np.float32) ** 2).sum(-1)
            mask = d2 < np.float32(cutoff) ** 2
            if not S.any():
                np.fill_diagonal(mask, False)
            i, j = np.nonzero(mask)
            i_all.append(i), j_all.append(j), s_all.append(np.broadcast_to(S, (len(i), 3)))
        i, j, S = np.concatenate(i_all), np.concatenate(j_all), np.concatenate(s_all)
        S = S + coef[i] - coef[j]  # shifts for the UNWRAPPED positions (neighbourhood_list_torch.py:85-88)
        order = np.lexsort((S[:, 2], S[:, 1], S[:, 0], j, i))
        eis.append(np.stack([idx[i[order]], idx[j[order]]]))
        shs.append(S[order])
    return (torch.tensor(np.concatenate(eis, axis=1), dtype=torch.long), torch.tensor(np.concatenate(shs), dtype=torch.float32))


def make_batch(name: str, rank: int = 0, n_mol: int | None = None, edges: bool = True):
    """-> (batch dict on CPU, targets dict, model kwargs (without avg_num_neighbours), mlp_irreps).
    ``edges=False`` leaves ``edge_index`` out (the caller builds it with the GPU neighbour list; mandatory for the
    water box, whose 2 M edges are not built on the host)."""
    cfg = CONFIGS[name]
    n_mol = n_mol or cfg["n_mol"]
    seed = cfg["seed"] + 1000 * rank
    data = {}
    if name == "water_box_infer":
        box = cfg["box"] * (n_mol / 8000.0) ** (1.0 / 3.0)
        pos, z, cell = water_box(n_mol, box, seed)
        batch = torch.zeros(pos.shape[0], dtype=torch.long)
        n_graphs = 1
        data["cell"] = cell
        edges = False
    elif name == "mace_large_bulk_train":
        pos, z, batch, cell = rocksalt_cells(n_mol, cfg["reps"], cfg["lattice"], cfg["sigma"], cfg["Z"], seed,
                                               avoid_cutoff=cfg["model"]["cut_off"])
        n_graphs = n_mol
        data["cell"] = cell
    else:
        # weak scaling = fixed work per GPU: every rank gets molecules of the SAME sizes as rank 0 (its own coordinates and
        # species), so the step time of N ranks is not the time of whichever rank drew the largest molecules
        sizes = None
        if rank != 0:
            _, _, b0 = random_molecules(n_mol, cfg["atoms"][0], cfg["atoms"][1], cfg["Z"], cfg["seed"], probs=cfg["probs"])
            sizes = torch.bincount(b0, minlength=n_mol).tolist()
        pos, z, batch = random_molecules(n_mol, cfg["atoms"][0], cfg["atoms"][1], cfg["Z"], seed, probs=cfg["probs"], sizes=sizes)
        n_graphs = n_mol
    data.update({
        "coordinates": pos,
        "node_attrs": torch.nn.functional.one_hot(z, cfg["Z"]).float(),
        "batch": batch,
        "num_graphs": torch.tensor(n_graphs),
    })
    if edges:
        if cfg["pbc"]:
            data["edge_index"], data["cell_shift_vector"] = radius_graph_pbc_small(pos, batch, data["cell"], cfg["model"]["cut_off"])
        else:
            data["edge_index"] = radius_graph_n2(pos, batch, cfg["model"]["cut_off"])
    g = torch.Generator().manual_seed(cfg["seed"] + 7 + rank)
    targets = {"energy": torch.randn(n_graphs, 1, generator=g), "forces": 0.1 * torch.randn(pos.shape[0], 3, generator=g)}
    return data, targets, dict(cfg["model"]), cfg["mlp"]


_DEGREE_CACHE: dict = {}


def mean_degree(name: str) -> float:
    """avg_num_neighbours is a model hyper-parameter: fixed from the rank-0 batch so that all ranks share weights"""
    if name not in _DEGREE_CACHE:
        cfg = CONFIGS[name]
        if "avg_num_neighbours" in cfg:
            _DEGREE_CACHE[name] = float(cfg["avg_num_neighbours"])
        else:
            data, _, _, _ = make_batch(name, 0)
            _DEGREE_CACHE[name] = data["edge_index"].shape[1] / data["coordinates"].shape[0]
    return _DEGREE_CACHE[name]


def build_model(name: str, oracle: bool = False, seed: int = 0):
    """The workload's model with the reference initialisers under ``torch.manual_seed(seed)``.  ``oracle=True`` builds
    the CPU restatement (tests / bench baseline legs only); both take the same state dict."""
    cfg = CONFIGS[name]
    kw = dict(cfg["model"])
    kw["avg_num_neighbours"] = mean_degree(name)
    torch.manual_seed(seed)
    if oracle:
        if cfg["family"] == "mace":
            from oracle.ref_model import RefPooledMACE as cls
        else:
            from oracle.ref_nequip import RefPooledNequip as cls
    elif cfg["family"] == "mace":
        from .mace import PooledMACEModule as cls
    else:
        from .nequip import PooledNequipModule as cls
    return cls(mlp_irreps=cfg["mlp"], **kw)


def energy_forces(model, data: dict, training: bool = False):
    """one energy + forces evaluation through the reference-shaped API (lightning/module.py:28-48, 65-95)"""
    from .mace import compute_forces_by_gradient

    d = dict(data)
    pos = d["coordinates"].detach().requires_grad_(True)
    d["coordinates"] = pos
    e = model(d)["y_graph_scalars"]
    return e, compute_forces_by_gradient(e, pos, training=training)


def training_loss(model, data: dict, targets: dict) -> torch.Tensor:
    """forward -> forces (create_graph) -> MSE(E) + MSE(F)   (reference lightning/module.py:65-95, mace_module.py:175-185)"""
    d = dict(data)
    pos = d["coordinates"].detach().requires_grad_(True)
    d["coordinates"] = pos
    e = model(d)["y_graph_scalars"]
    from . import ops

    with ops.input_grads_only():  # the force pass: gradients w.r.t. the coordinates only
        (g,) = torch.autograd.grad([e], [pos], grad_outputs=[torch.ones_like(e)], retain_graph=True, create_graph=True)
    if e.is_cuda:  # MSE(E) + MSE(F) in one kernel; the CPU legs (oracle / reference arm) use the stock losses
        return ops.energy_force_mse(e, targets["energy"], g, targets["forces"], 1.0, 1.0)
    return torch.nn.functional.mse_loss(e, targets["energy"]) + torch.nn.functional.mse_loss(-g, targets["forces"])


def make_optimizer(model, capturable: bool = False, grad_scale: float = 1.0):
    """reference default: Adam(lr=1e-2, amsgrad=True, weight_decay=5e-7) (mace/supervised/default_configs.py:25-33) over the
    reference's one-group-per-parameter list for MACE (mace_module.py:131-157, optim.reference_param_groups), a single
    group for NequIP (base-class configure_optimizers).  CUDA parameters get the fused single-kernel implementation of
    the same update rule (physicsml_b200.optim.FusedAdam, always graph-capturable); CPU parameters (the oracle /
    reference arm) get torch.optim.Adam."""
    from .optim import FusedAdam, reference_param_groups

    if hasattr(model, "mace"):
        groups = [g for g in reference_param_groups(model, 5e-7) if g["params"][0].requires_grad]
        wd = 0.0
    else:
        groups = [p for p in model.parameters() if p.requires_grad and p.numel() > 0]
        wd = 5e-7
    first = groups[0]["params"][0] if isinstance(groups[0], dict) else groups[0]
    if first.is_cuda:
        return FusedAdam(groups, lr=1e-2, weight_decay=wd, amsgrad=True, grad_scale=grad_scale)
    return torch.optim.Adam(groups, lr=1e-2, amsgrad=True, weight_decay=wd, capturable=capturable)
