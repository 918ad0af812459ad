"""CUDA path against the golden energies / forces produced by the reference's own unmodified MACE source
(tests/golden/make_golden.py).  Bar: energies 1e-5 relative, forces 1e-4 absolute (BASELINE.json north_star)."""
import glob
import os

import pytest
import torch

import helpers  # noqa: F401

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
MACE_CASES = sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(GOLD, "mace_*.pt")))
NEQUIP_CASES = sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(GOLD, "nequip_*.pt")))


@pytest.mark.parametrize("name", MACE_CASES)
def test_cuda_reproduces_reference_golden(cuda, name):
    from physicsml_b200.mace import PooledMACEModule, compute_forces_by_gradient

    g = torch.load(os.path.join(GOLD, name + ".pt"), weights_only=False)
    model = PooledMACEModule(mlp_irreps=g["mlp"], **g["kw"])
    res = model.load_state_dict(g["state_dict"], strict=False)
    assert not res.unexpected_keys
    model = model.to(cuda)
    Z = g["kw"]["num_node_feats"]
    data = {"coordinates": g["coordinates"].to(cuda).requires_grad_(True), "edge_index": g["edge_index"].to(cuda),
            "batch": g["batch"].to(cuda), "node_attrs": torch.nn.functional.one_hot(g["z"], Z).float().to(cuda),
            "num_graphs": torch.tensor(int(g["batch"].max()) + 1)}
    e = model(data)["y_graph_scalars"]
    f = compute_forces_by_gradient(e, data["coordinates"])
    assert (e.cpu() - g["energy"]).abs().max() <= 1e-5 * g["energy"].abs().max()
    assert (f.cpu() - g["forces"]).abs().max() <= 1e-4


@pytest.mark.parametrize("name", NEQUIP_CASES)
def test_cuda_nequip_reproduces_reference_golden(cuda, name):
    from physicsml_b200.mace import compute_forces_by_gradient
    from physicsml_b200.nequip import PooledNequipModule

    g = torch.load(os.path.join(GOLD, name + ".pt"), weights_only=False)
    model = PooledNequipModule(mlp_irreps=g["mlp"], **g["kw"])
    res = model.load_state_dict(g["state_dict"], strict=False)
    assert not res.unexpected_keys
    assert [bool(l.resnet) for l in model.nequip.conv_layers] == g["resnet_layers"]
    model = model.to(cuda).eval()
    Z = g["kw"]["num_node_feats"]
    data = {"coordinates": g["coordinates"].to(cuda).requires_grad_(True), "edge_index": g["edge_index"].to(cuda),
            "batch": g["batch"].to(cuda), "node_attrs": torch.nn.functional.one_hot(g["z"], Z).float().to(cuda),
            "num_graphs": torch.tensor(int(g["batch"].max()) + 1)}
    e = model(data)["y_graph_scalars"]
    f = compute_forces_by_gradient(e, data["coordinates"])
    assert (e.cpu() - g["energy"]).abs().max() <= 1e-5 * g["energy"].abs().max()
    assert (f.cpu() - g["forces"]).abs().max() <= 1e-4
