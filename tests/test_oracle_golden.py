"""The standalone oracle (oracle/ref_model.py, oracle/ref_nl.py) against golden vectors produced by the reference's own
unmodified source files (tests/golden/make_golden.py).  Runs anywhere (CPU)."""
import glob
import os

import pytest
import torch

import helpers  # noqa: F401

GOLD = os.path.join(os.path.dirname(__file__), "golden")
MACE_CASES = sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(GOLD, "mace_*.pt")))
NEQUIP_CASES = sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(GOLD, "nequip_*.pt")))
NL_CASES = sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(GOLD, "nl_*.pt")))


@pytest.mark.parametrize("name", MACE_CASES)
def test_oracle_reproduces_reference_energy_and_forces(name):
    from oracle.ref_model import RefPooledMACE, energy_and_forces

    g = torch.load(os.path.join(GOLD, name + ".pt"), weights_only=False)
    model = RefPooledMACE(mlp_irreps=g["mlp"], **g["kw"])
    missing = model.load_state_dict(g["state_dict"], strict=False)
    assert not missing.unexpected_keys
    assert all(("U_matrices" in k) or ("_w3j_" in k) or ("output_mask" in k) for k in missing.missing_keys)
    Z = g["kw"]["num_node_feats"]
    data = {"coordinates": g["coordinates"], "edge_index": g["edge_index"], "batch": g["batch"],
            "node_attrs": torch.nn.functional.one_hot(g["z"], Z).float(), "num_graphs": torch.tensor(int(g["batch"].max()) + 1)}
    e, f = energy_and_forces(model, data)
    assert (e - g["energy"]).abs().max() <= 1e-6 * g["energy"].abs().max()
    assert (f - g["forces"]).abs().max() <= 1e-6


@pytest.mark.parametrize("name", NEQUIP_CASES)
def test_oracle_reproduces_reference_nequip(name):
    """oracle/ref_nequip.py against the unmodified reference nequip.py (over the shims), incl. the fixture checkpoint"""
    from oracle.ref_model import energy_and_forces
    from oracle.ref_nequip import RefPooledNequip

    g = torch.load(os.path.join(GOLD, name + ".pt"), weights_only=False)
    model = RefPooledNequip(mlp_irreps=g["mlp"], **g["kw"])
    missing = model.load_state_dict(g["state_dict"], strict=False)
    assert not missing.unexpected_keys
    assert all(("_w3j_" in k) or ("output_mask" in k) for k in missing.missing_keys)
    assert [bool(l.resnet) for l in model.nequip.conv_layers] == g["resnet_layers"]
    Z = g["kw"]["num_node_feats"]
    data = {"coordinates": g["coordinates"], "edge_index": g["edge_index"], "batch": g["batch"],
            "node_attrs": torch.nn.functional.one_hot(g["z"], Z).float(), "num_graphs": torch.tensor(int(g["batch"].max()) + 1)}
    e, f = energy_and_forces(model, data)
    assert (e - g["energy"]).abs().max() <= 1e-6 * g["energy"].abs().max()
    assert (f - g["forces"]).abs().max() <= 1e-6


@pytest.mark.parametrize("name", NL_CASES)
def test_oracle_neighbour_list_matches_reference(name):
    from oracle.ref_nl import ref_radius_graph

    g = torch.load(os.path.join(GOLD, name + ".pt"), weights_only=False)
    ei, sh = ref_radius_graph(g["pos"], g["cutoff"], g["pbc"], g["cell"])
    assert ei.shape == g["edge_index"].shape
    assert torch.equal(ei, g["edge_index"].long())
    assert torch.equal(sh, g["shift"].to(sh.dtype))


def test_reference_known_answer_edge_counts():
    """20 / 12 edges for CH4 / NH3 and 50 edges, 16 nodes for the first four gdb9 molecules: the only neighbour-list
    known answers in the reference's test-suite (tests/rdkit/core/data/test_graph_dataset.py:42,50; test_datamodule.py:28-35)"""
    from oracle.ref_nl import ref_radius_graph

    g = torch.load(os.path.join(GOLD, "gdb9_edges.pt"), weights_only=False)
    counts = [ref_radius_graph(p, 5.0)[0].shape[1] for p in g["positions"]]
    assert counts == g["edge_counts"]
    assert counts[0] == 20 and counts[1] == 12 and sum(counts[:4]) == 50
    assert sum(p.shape[0] for p in g["positions"][:4]) == 16
