"""End-to-end parity of the CUDA NequIP path (reference-shaped module API, shared fused edge kernels + gate kernel)
against the oracle restatement (oracle/ref_nequip.py, itself pinned to the unmodified reference by the goldens) on
identical inputs and weights.  Bar (BASELINE.json north_star): fp32 energies within 1e-5 relative, forces within 1e-4
absolute."""
import pytest
import torch

import helpers

pytestmark = pytest.mark.gpu

BASE = dict(cut_off=4.5, num_bessel=8, bessel_basis_trainable=False, num_polynomial_cutoff=6, self_connection=True, resnet=True,
            num_edge_feats=0, avg_num_neighbours=11.0)
CONFIGS = {
    "f8_l2_4layers": dict(mlp="16x0e", Z=5, kw=dict(BASE, num_layers=4, max_ell=2, parity=True, num_features=8, num_node_feats=5)),
    "f64_l2_4layers": dict(mlp="16x0e", Z=5, kw=dict(BASE, num_layers=4, max_ell=2, parity=True, num_features=64, num_node_feats=5)),
    "f32_l1_3layers": dict(mlp="8x0e", Z=3, kw=dict(BASE, num_layers=3, max_ell=1, parity=True, num_features=32, num_node_feats=3)),
    "f16_l2_even": dict(mlp="8x0e", Z=3, kw=dict(BASE, num_layers=3, max_ell=2, parity=False, num_features=16, num_node_feats=3,
                                                  self_connection=False)),
}


def _models(name, device):
    from oracle.ref_nequip import RefPooledNequip
    from physicsml_b200.nequip import PooledNequipModule

    cfg = CONFIGS[name]
    torch.manual_seed(0)
    ref = RefPooledNequip(mlp_irreps=cfg["mlp"], **cfg["kw"])
    mine = PooledNequipModule(mlp_irreps=cfg["mlp"], **cfg["kw"])
    res = mine.load_state_dict({k: v for k, v in ref.state_dict().items()}, strict=False)
    assert not res.unexpected_keys and all("_w3j_" in k or "output_mask" in k for k in res.missing_keys)
    return ref, mine.to(device), cfg


@pytest.mark.parametrize("name", list(CONFIGS))
def test_energy_forces_parity(cuda, name):
    from oracle.ref_model import energy_and_forces
    from physicsml_b200.mace import compute_forces_by_gradient

    ref, mine, cfg = _models(name, cuda)
    pos, z, batch = helpers.random_molecules(6 if "f64" in name else 10, 10, 22, cfg["Z"], seed=303)
    data = helpers.make_batch(pos, z, batch, cfg["Z"], cfg["kw"]["cut_off"])
    e_ref, f_ref = energy_and_forces(ref, data)
    d = helpers.to_device(data, cuda)
    d["coordinates"].requires_grad_(True)
    e = mine(d)["y_graph_scalars"]
    f = compute_forces_by_gradient(e, d["coordinates"])
    e, f = e.detach().cpu(), f.detach().cpu()
    scale = e_ref.abs().max().item()
    print(name, "E rel err", (e - e_ref).abs().max().item() / scale, "F abs err", (f - f_ref).abs().max().item(), "|F|max", f_ref.abs().max().item())
    assert (e - e_ref).abs().max().item() <= 1e-5 * scale
    assert (f - f_ref).abs().max().item() <= 1e-4
    # the batch-dict contract of the reference (nequip.py:101,119-122): these keys are written for downstream heads
    for k in ("node_feats", "edge_feats", "edge_attrs"):
        assert k in d


@pytest.mark.parametrize("scal,gates,gated", [("8x0o+8x0e", "8x0e+8x0e+8x0e+8x0e", "8x1o+8x1e+8x2o+8x2e"), ("4x0e", "4x0e+4x0e", "4x1o+4x2e"),
                                              ("64x0e", "", "")])
def test_gate_op(cuda, scal, gates, gated):
    """value, cotangent map and its derivative (force-matching training) against the oracle Gate under autograd"""
    from e3nn import o3
    from oracle.ref_nequip import _Gate
    from physicsml_b200 import ops
    from physicsml_b200.plans import GatePlan

    ref = _Gate(o3.Irreps(scal), o3.Irreps(gates), o3.Irreps(gated))
    plan = GatePlan(scal, gates, gated)
    g = torch.Generator().manual_seed(4)
    N = 77
    x0 = torch.randn(N, plan.d_in, generator=g)
    w0 = torch.randn(N, plan.d_out, generator=g)
    v0 = torch.randn(N, plan.d_in, generator=g)

    def run(fn, dev):
        x = x0.to(dev).requires_grad_(True)
        w = w0.to(dev).requires_grad_(True)
        y = fn(x)
        (dx,) = torch.autograd.grad([y], [x], [w], create_graph=True)
        ddx, ddw = torch.autograd.grad([(dx * v0.to(dev)).sum()], [x, w])
        return [t.detach().cpu() for t in (y, dx, ddx, ddw)]

    a = run(lambda x: ops.gate(x, plan.handle), cuda)
    b = run(ref, "cpu")
    for name, p, q in zip(("y", "dx", "ddx", "ddy"), a, b):
        assert (p - q).abs().max() <= 2e-5 * max(1.0, q.abs().max().item()), name


def test_training_step_gradients(cuda):
    """force-matching loss through the double-backward formulas (shared edge kernels + gate_bwd2) vs the oracle's autograd"""
    ref, mine, cfg = _models("f8_l2_4layers", cuda)
    pos, z, batch = helpers.random_molecules(4, 8, 14, cfg["Z"], seed=202)
    data = helpers.make_batch(pos, z, batch, cfg["Z"], cfg["kw"]["cut_off"])
    g = torch.Generator().manual_seed(9)
    e_t = torch.randn(4, 1, generator=g)
    f_t = torch.randn(pos.shape[0], 3, generator=g) * 0.1

    def loss_of(model, d, dev):
        d = dict(d)
        d["coordinates"] = d["coordinates"].clone().requires_grad_(True)
        e = model(d)["y_graph_scalars"]
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
