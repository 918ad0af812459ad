"""Generates the committed golden fixtures by running the reference's OWN, UNMODIFIED source files (imported from
/root/reference over the oracle's e3nn / torch_geometric / opt_einsum_fx restatements; see oracle/import_reference.py).

Only runs in the build container (needs /root/reference).  Outputs (all small, fp32):
  fixture_cg.pt      Wigner-3j buffers and U_matrices copied out of the reference's own test checkpoints
                     (tests/data/{mace,nequip}_model/model_artefacts/module_checkpoint.ckpt) — the e3nn-convention pins
  mace_*.pt          inputs, state dict, energies and autograd forces of the reference MACE (models/mace/modules/mace.py)
                     + PooledReadoutHead on seeded synthetic molecules
  nequip_*.pt        the same for the reference NequIP (models/nequip/modules/nequip.py) + PooledReadoutHead
  nl_*.pt            inputs and sorted edge sets of the reference construct_edge_indices_and_attrs
                     (lightning/graph_datasets/neighbourhood_list_torch.py:53)
  bondmerge_nl.pt   inputs and outputs of the same function's bond-feature merge branch (:98-156), open and periodic
  gdb9_edges.pt      first six molecules of the reference's tests/data/gdb9.sdf with the edge counts its tests assert
                     (tests/rdkit/core/data/test_graph_dataset.py:42,50; test_datamodule.py:28-35)

    python tests/golden/make_golden.py
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import helpers  # noqa: E402
from oracle import import_reference as ir  # noqa: E402

assert ir.available(), "/root/reference is required to (re)generate the golden fixtures"
mace_mod = ir.load("physicsml.models.mace.modules.mace")
nequip_mod = ir.load("physicsml.models.nequip.modules.nequip")
nl_mod = ir.load("physicsml.lightning.graph_datasets.neighbourhood_list_torch")
from e3nn import o3  # noqa: E402  (oracle restatement)

from oracle.ref_nl import sorted_edge_set  # noqa: E402


class RefPooled(torch.nn.Module):
    """backbone + graph-scalar head exactly as PooledMACEModule wires them (mace/supervised/mace_module.py:30-70,96-108)"""

    def __init__(self, mlp_irreps, **kw):
        super().__init__()
        self.mace = mace_mod.MACE(**kw)
        self.graph_scalars_head = mace_mod.PooledReadoutHead(
            list_in_irreps=self.mace.out_irreps, mlp_irreps=o3.Irreps(mlp_irreps), out_irreps=o3.Irreps("1x0e"),
            scaling_std=1.0, scaling_mean=0.0)

    def forward(self, data):
        return {"y_graph_scalars": self.graph_scalars_head(self.mace(data))}


class RefPooledNequip(torch.nn.Module):
    """backbone + graph-scalar head as PooledNequipModule wires them (nequip/supervised/nequip_module.py:37-80,102-113)"""

    def __init__(self, mlp_irreps, **kw):
        super().__init__()
        self.nequip = nequip_mod.Nequip(**kw)
        self.graph_scalars_head = nequip_mod.PooledReadoutHead(
            irrreps_in=self.nequip.out_irreps[-1], mlp_irreps=o3.Irreps(mlp_irreps), out_irreps=o3.Irreps("1x0e"),
            scaling_std=1.0, scaling_mean=0.0)

    def forward(self, data):
        return {"y_graph_scalars": self.graph_scalars_head(self.nequip(data))}


def energy_forces(model, data):
    data = dict(data)
    pos = data["coordinates"].clone().requires_grad_(True)
    data["coordinates"] = pos
    e = model(data)["y_graph_scalars"]
    (g,) = torch.autograd.grad([e], [pos], [torch.ones_like(e)])
    return e.detach(), (-g).detach()


def main():
    # ---- 1. CG pins from the reference's checkpoints
    pins = {}
    for name in ("mace", "nequip"):
        sd = torch.load(f"/root/reference/tests/data/{name}_model/model_artefacts/module_checkpoint.ckpt", map_location="cpu",
                        weights_only=False)["state_dict"]
        for k, v in sd.items():
            if "_w3j_" in k:
                pins.setdefault("w3j", {})[k.split("_w3j_")[1]] = v.clone()
            if "U_matrices" in k and name == "mace" and "messages.0" in k:
                c = k.split("contractions.")[-1]
                pins.setdefault("U", {})[c] = v.clone()
    torch.save(pins, os.path.join(HERE, "fixture_cg.pt"))
    print("fixture_cg.pt:", {k: len(v) for k, v in pins.items()})

    # ---- 2. reference MACE energies / forces
    cases = {
        "mace_fixture_ckpt": dict(mlp="5x0e", Z=4, ckpt=True, kw=dict(
            cut_off=5.0, num_bessel=8, num_polynomial_cutoff=6, max_ell=2, num_interactions=2, num_node_feats=4, num_edge_feats=0,
            hidden_irreps="4x0e + 4x1o", avg_num_neighbours=10.0, correlation=2)),
        "mace_small_c16": dict(mlp="16x0e", Z=5, ckpt=False, kw=dict(
            cut_off=5.0, num_bessel=8, num_polynomial_cutoff=5, max_ell=2, num_interactions=2, num_node_feats=5, num_edge_feats=0,
            hidden_irreps="16x0e + 16x1o", avg_num_neighbours=14.0, correlation=3)),
        "mace_lmax3_c8": dict(mlp="16x0e", Z=3, ckpt=False, kw=dict(
            cut_off=4.5, num_bessel=8, num_polynomial_cutoff=5, max_ell=3, num_interactions=2, num_node_feats=3, num_edge_feats=0,
            hidden_irreps="8x0e + 8x1o", avg_num_neighbours=12.0, correlation=3)),
        "mace_l2_c8_3layers": dict(mlp="8x0e", Z=3, ckpt=False, kw=dict(
            cut_off=4.5, num_bessel=8, num_polynomial_cutoff=5, max_ell=3, num_interactions=3, num_node_feats=3, num_edge_feats=0,
            hidden_irreps="8x0e + 8x1o + 8x2e", avg_num_neighbours=12.0, correlation=3)),
    }
    for name, c in cases.items():
        torch.manual_seed(0)
        model = RefPooled(c["mlp"], **c["kw"])
        if c["ckpt"]:
            sd = torch.load("/root/reference/tests/data/mace_model/model_artefacts/module_checkpoint.ckpt", map_location="cpu",
                            weights_only=False)["state_dict"]
            model.load_state_dict(sd, strict=True)
        pos, z, batch = helpers.random_molecules(5, 8, 16, c["Z"], seed=101)
        data = helpers.make_batch(pos, z, batch, c["Z"], c["kw"]["cut_off"])
        e, f = energy_forces(model, data)
        sd = {k: v.clone() for k, v in model.state_dict().items() if "U_matrices" not in k and "_w3j_" not in k and "output_mask" not in k}
        torch.save({"kw": c["kw"], "mlp": c["mlp"], "state_dict": sd, "coordinates": pos, "z": z, "batch": batch,
                    "edge_index": data["edge_index"], "energy": e, "forces": f}, os.path.join(HERE, name + ".pt"))
        print(name, "E", e.flatten().tolist()[:3], "|F|max", f.abs().max().item(), "edges", data["edge_index"].shape[1])

    # ---- 2b. reference NequIP energies / forces
    ncases = {
        "nequip_fixture_ckpt": dict(mlp="8x0e", Z=4, ckpt=True, kw=dict(
            cut_off=5.0, num_layers=2, max_ell=2, parity=True, num_features=4, num_bessel=8, bessel_basis_trainable=True,
            num_polynomial_cutoff=6, self_connection=True, resnet=True, num_node_feats=4, num_edge_feats=0, avg_num_neighbours=10.0)),
        "nequip_l2_f8_4layers": dict(mlp="16x0e", Z=5, ckpt=False, kw=dict(
            cut_off=4.5, num_layers=4, max_ell=2, parity=True, num_features=8, num_bessel=8, bessel_basis_trainable=False,
            num_polynomial_cutoff=6, self_connection=True, resnet=True, num_node_feats=5, num_edge_feats=0, avg_num_neighbours=11.0)),
        "nequip_l1_f16_noparity": dict(mlp="8x0e", Z=3, ckpt=False, kw=dict(
            cut_off=4.0, num_layers=3, max_ell=1, parity=False, num_features=16, num_bessel=8, bessel_basis_trainable=False,
            num_polynomial_cutoff=6, self_connection=False, resnet=False, num_node_feats=3, num_edge_feats=0, avg_num_neighbours=8.0)),
    }
    for name, c in ncases.items():
        torch.manual_seed(0)
        model = RefPooledNequip(c["mlp"], **c["kw"])
        if c["ckpt"]:
            sd = torch.load("/root/reference/tests/data/nequip_model/model_artefacts/module_checkpoint.ckpt", map_location="cpu",
                            weights_only=False)["state_dict"]
            model.load_state_dict(sd, strict=True)
        pos, z, batch = helpers.random_molecules(5, 8, 16, c["Z"], seed=303)
        data = helpers.make_batch(pos, z, batch, c["Z"], c["kw"]["cut_off"])
        e, f = energy_forces(model, data)
        sd = {k: v.clone() for k, v in model.state_dict().items() if "_w3j_" not in k and "output_mask" not in k}
        torch.save({"kw": c["kw"], "mlp": c["mlp"], "state_dict": sd, "coordinates": pos, "z": z, "batch": batch,
                    "edge_index": data["edge_index"], "energy": e, "forces": f, "resnet_layers": [bool(l.resnet) for l in model.nequip.conv_layers]},
                   os.path.join(HERE, name + ".pt"))
        print(name, "E", e.flatten().tolist()[:3], "|F|max", f.abs().max().item(), "edges", data["edge_index"].shape[1],
              "resnet", [bool(l.resnet) for l in model.nequip.conv_layers])

    # ---- 3. reference neighbour lists
    g = torch.Generator().manual_seed(404)
    tri = torch.tensor([[10.0, 0, 0], [3.0, 9.0, 0], [-2.0, 1.5, 8.0]])
    nl_cases = {
        "nl_small_cell": (torch.rand(5, 3, generator=g) * 4, 5.0, (True, True, True), torch.eye(3) * 4),
        "nl_cubic_unwrapped": (torch.rand(150, 3, generator=g) * 12 + torch.randn(150, 3, generator=g) * 4, 5.0, (True, True, True), torch.eye(3) * 12),
        "nl_triclinic": (torch.rand(120, 3, generator=g) * 10 - 3, 4.5, (True, True, True), tri),
        "nl_open": (torch.randn(40, 3, generator=g) * 2.5, 5.0, None, None),
    }
    for name, (pos, rc, pbc, cell) in nl_cases.items():
        ei, _, sh = nl_mod.construct_edge_indices_and_attrs(pos, rc, None, None, pbc, cell, False)
        assert bool((ei[0][1:] >= ei[0][:-1]).all()), "reference output is sorted by sender"
        ei_s, sh_s = sorted_edge_set(ei, sh)
        torch.save({"pos": pos, "cutoff": rc, "pbc": pbc, "cell": cell, "edge_index": ei_s.to(torch.int32), "shift": sh_s.to(torch.int8)},
                   os.path.join(HERE, name + ".pt"))
        print(name, "E", ei.shape[1])

    # ---- 4. the edge counts the reference's own tests assert, from its SDF fixture (parsed as text; no RDKit)
    mols, cur = [], []
    with open("/root/reference/tests/data/gdb9.sdf") as f:
        lines = f.read().split("$$$$\n")
    for block in lines[:6]:
        rows = block.split("\n")
        n_at = int(rows[3][:3])
        xyz = [[float(x) for x in r.split()[:3]] for r in rows[4 : 4 + n_at]]
        mols.append(torch.tensor(xyz))
    counts = []
    for p in mols:
        ei, _, _ = nl_mod.construct_edge_indices_and_attrs(p, 5.0, None, None, None, None, False)
        counts.append(ei.shape[1])
    assert counts[:2] == [20, 12] and sum(counts[:4]) == 50, counts
    torch.save({"positions": mols, "edge_counts": counts}, os.path.join(HERE, "gdb9_edges.pt"))
    print("gdb9 edge counts", counts)
    bond_merge()


def bond_merge():
    """---- 5. the bond-feature merge branch of construct_edge_indices_and_attrs (neighbourhood_list_torch.py:98-156):
    supplied edges + attributes merged into the radius graph, open and periodic"""
    pos, _, _ = helpers.random_molecules(1, 9, 9, 3, seed=5)
    bonds = torch.tensor([[0, 1], [1, 2], [2, 3], [3, 4], [0, 5]])
    attrs = torch.tensor([[1.0, 0.0], [0.0, 1.0], [1.0, 0.0], [0.5, 0.5], [0.0, 2.0]])
    out = {"pos": pos, "bonds": bonds, "attrs": attrs, "cutoff": 5.0}
    ei, ea, sh = nl_mod.construct_edge_indices_and_attrs(pos, 5.0, bonds, attrs, None, None, False)
    out["open"] = {"edge_index": ei, "edge_attrs": ea, "shift": sh}
    cell = torch.eye(3) * 14.0
    ei, ea, sh = nl_mod.construct_edge_indices_and_attrs(pos + 6.0, 5.0, bonds, attrs, (True, True, True), cell, False)
    out["pbc"] = {"cell": cell, "offset": 6.0, "edge_index": ei, "edge_attrs": ea, "shift": sh}
    torch.save(out, os.path.join(HERE, "bondmerge_nl.pt"))
    print("bondmerge_nl: E", ei.shape[1], "attr sum", ea.sum(0).tolist())


if __name__ == "__main__":
    main()
