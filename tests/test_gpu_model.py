"""End-to-end parity of the CUDA MACE path (through the reference-shaped module API) against the oracle restatement
on identical inputs and weights.  Bar (BASELINE.json north_star): fp32 energies within 1e-5 relative, forces within
1e-4 absolute (eV/A)."""
import pytest
import torch

import helpers

pytestmark = pytest.mark.gpu

CONFIGS = {
    "fixture": dict(mlp="5x0e", Z=4, kw=dict(cut_off=5.0, num_bessel=8, num_polynomial_cutoff=6, max_ell=2, num_interactions=2,
                    num_node_feats=4, num_edge_feats=0, hidden_irreps="4x0e + 4x1o", avg_num_neighbours=10.0, correlation=2)),
    "small": dict(mlp="16x0e", Z=5, kw=dict(cut_off=5.0, num_bessel=8, num_polynomial_cutoff=5, max_ell=2, num_interactions=2,
                  num_node_feats=5, num_edge_feats=0, hidden_irreps="32x0e + 32x1o", avg_num_neighbours=15.0, correlation=3)),
    "medium": dict(mlp="16x0e", Z=10, kw=dict(cut_off=5.0, num_bessel=8, num_polynomial_cutoff=5, max_ell=3, num_interactions=2,
                   num_node_feats=10, num_edge_feats=0, hidden_irreps="128x0e + 128x1o", avg_num_neighbours=20.0, correlation=3)),
    # BASELINE configs[4] (MACE large: 256 channels, L = 2, correlation 3) at reduced graph size: exercises the 17-path
    # tensor product (DOUT = 71: register-streaming K4), the L <= 2 symmetric contraction and 256-wide element slabs
    "large": dict(mlp="16x0e", Z=4, kw=dict(cut_off=4.0, num_bessel=8, num_polynomial_cutoff=5, max_ell=3, num_interactions=2,
                  num_node_feats=4, num_edge_feats=0, hidden_irreps="256x0e + 256x1o + 256x2e", avg_num_neighbours=12.0,
                  correlation=3)),
    "l2": dict(mlp="16x0e", Z=4, kw=dict(cut_off=4.5, num_bessel=8, num_polynomial_cutoff=5, max_ell=3, num_interactions=3,
               num_node_feats=4, num_edge_feats=0, hidden_irreps="16x0e + 16x1o + 16x2e", avg_num_neighbours=12.0, correlation=3)),
}


def _models(name, device):
    from oracle.ref_model import RefPooledMACE
    from physicsml_b200.mace import PooledMACEModule

    cfg = CONFIGS[name]
    torch.manual_seed(0)
    ref = RefPooledMACE(mlp_irreps=cfg["mlp"], **cfg["kw"])
    mine = PooledMACEModule(mlp_irreps=cfg["mlp"], **cfg["kw"])
    mine.load_state_dict(ref.state_dict(), strict=True)
    return ref, mine.to(device), cfg


@pytest.mark.parametrize("name", list(CONFIGS))
def test_energy_forces_parity(cuda, name):
    from oracle.ref_model import energy_and_forces
    from physicsml_b200.mace import compute_forces_by_gradient

    ref, mine, cfg = _models(name, cuda)
    n_mol = {"medium": 6, "large": 3}.get(name, 10)
    pos, z, batch = helpers.random_molecules(n_mol, 10, 14 if name == "large" else 22, cfg["Z"], seed=101)
    data = helpers.make_batch(pos, z, batch, cfg["Z"], cfg["kw"]["cut_off"])
    e_ref, f_ref = energy_and_forces(ref, data)
    e64, f64 = energy_and_forces(ref.double(), helpers.to_device(data, "cpu", torch.float64))
    d = helpers.to_device(data, cuda)
    d["coordinates"].requires_grad_(True)
    e = mine(d)["y_graph_scalars"]
    f = compute_forces_by_gradient(e, d["coordinates"])
    e, f = e.detach().cpu(), f.detach().cpu()
    scale = e64.abs().max().item()
    print(name, "E err vs fp32 oracle", (e - e_ref).abs().max().item() / scale, "vs fp64", (e - e64.float()).abs().max().item() / scale,
          "oracle32 vs 64", (e_ref - e64.float()).abs().max().item() / scale)
    print(name, "F err vs fp32 oracle", (f - f_ref).abs().max().item(), "vs fp64", (f - f64.float()).abs().max().item(),
          "|F|max", f64.abs().max().item())
    assert (e - e_ref).abs().max().item() <= 1e-5 * scale
    assert (f - f_ref).abs().max().item() <= 1e-4


@pytest.mark.parametrize("prune", [False, True])
def test_training_step_gradients(cuda, prune):
    """force-matching loss: d(loss)/d(params) through the double-backward formulas vs the oracle's autograd.
    prune=True runs the force pass under ops.input_grads_only() (what compute_forces_by_gradient and the benchmark's
    training step do): no weight gradients are computed in that pass, the parameter gradients of the loss must not
    change."""
    import contextlib

    from oracle.ref_model import RefPooledMACE  # noqa: F401
    from physicsml_b200 import ops
    from physicsml_b200.mace import compute_forces_by_gradient

    ref, mine, cfg = _models("small", cuda)
    pos, z, batch = helpers.random_molecules(4, 8, 14, cfg["Z"], seed=202)
    data = helpers.make_batch(pos, z, batch, cfg["Z"], cfg["kw"]["cut_off"])
    g = torch.Generator().manual_seed(9)
    e_t = torch.randn(4, 1, generator=g)
    f_t = torch.randn(pos.shape[0], 3, generator=g) * 0.1

    def loss_of(model, d, dev):
        d = dict(d)
        d["coordinates"] = d["coordinates"].clone().requires_grad_(True)
        e = model(d)["y_graph_scalars"]
        with (ops.input_grads_only() if (prune and dev != "cpu") else contextlib.nullcontext()):
            (gr,) = torch.autograd.grad([e], [d["coordinates"]], [torch.ones_like(e)], create_graph=True, retain_graph=True)
        return ((e - e_t.to(dev)) ** 2).mean() + ((-gr - f_t.to(dev)) ** 2).mean()

    l_ref = loss_of(ref, data, "cpu")
    l_ref.backward()
    l = loss_of(mine, helpers.to_device(data, cuda), cuda)
    l.backward()
    assert abs(l.item() - l_ref.item()) <= 1e-5 * abs(l_ref.item())
    pr = dict(ref.named_parameters())
    worst = 0.0
    for k, p in mine.named_parameters():
        if not p.requires_grad or p.numel() == 0:
            continue
        a, b = p.grad, pr[k].grad
        assert (a is None) == (b is None), k
        if a is None:
            continue
        err = (a.cpu() - b).abs().max().item() / max(b.abs().max().item(), 1e-12)
        worst = max(worst, err)
        assert err < 2e-4, (k, err)
    print("worst relative parameter-gradient error", worst)


def test_edge_attribute_inputs(cuda):
    """num_edge_feats > 0 (reference mace.py:151-152): per-edge attributes (bond features merged into the
    radius graph by construct_edge_indices_and_attrs) are concatenated to the Bessel features in front of the radial MLP"""
    from oracle.ref_model import RefPooledMACE, energy_and_forces
    from physicsml_b200.mace import PooledMACEModule, compute_forces_by_gradient
    from physicsml_b200.neighbour_list import construct_edge_indices_and_attrs

    kw = dict(CONFIGS["small"]["kw"], num_edge_feats=2)
    torch.manual_seed(0)
    ref = RefPooledMACE(mlp_irreps="16x0e", **kw)
    mine = PooledMACEModule(mlp_irreps="16x0e", **kw)
    mine.load_state_dict(ref.state_dict(), strict=True)
    mine = mine.to(cuda)
    pos, z, batch = helpers.random_molecules(1, 14, 14, 5, seed=77)
    bonds = torch.tensor([[0, 1], [1, 2], [2, 3], [4, 5], [6, 7]])
    d = (pos[bonds[:, 0]] - pos[bonds[:, 1]]).norm(dim=1)
    bonds = bonds[d < 4.5]
    attrs = torch.tensor([[1.0, 0.0], [0.0, 1.0], [1.0, 1.0], [0.5, 0.0], [0.0, 2.0]])[d < 4.5]
    ei, ea, sh = construct_edge_indices_and_attrs(pos.to(cuda), 5.0, bonds.to(cuda), attrs.to(cuda), None, None, False)
    assert ea.shape == (ei.shape[1], 2) and float(ea.sum()) == 2 * float(attrs.sum())
    data = {"coordinates": pos, "edge_index": ei.cpu(), "edge_attrs": ea.cpu(), "node_attrs": torch.nn.functional.one_hot(z, 5).float(),
            "batch": batch, "num_graphs": torch.tensor(1)}
    e_ref, f_ref = energy_and_forces(ref, data)
    dd = helpers.to_device(data, cuda)
    dd["coordinates"].requires_grad_(True)
    e = mine(dd)["y_graph_scalars"]
    f = compute_forces_by_gradient(e, dd["coordinates"])
    assert (e.cpu() - e_ref).abs().max().item() <= 1e-5 * e_ref.abs().max().item()
    assert (f.cpu() - f_ref).abs().max().item() <= 1e-4
    # the attributes matter: zeroing them changes the energy
    dd2 = dict(dd, edge_attrs=torch.zeros_like(dd["edge_attrs"]))
    assert (mine(dd2)["y_graph_scalars"] - e).abs().max().item() > 1e-6
