"""GPU neighbour list (K1, through the C ABI) against the oracle restatement of the reference's semantics and the
golden edge sets produced by the reference itself.  Bar: the edge set is bit-exact after sorting."""
import glob
import os

import pytest
import torch

import helpers

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
NL_CASES = sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(GOLD, "nl_*.pt")))


def _check_plan(ei, plan, N):
    """the receiver-grouped permutation really groups by receiver, shares the sender row pointer, and is a bijection"""
    E = ei.shape[1]
    rowptr = plan.rowptr.long().cpu()
    perm = plan.perm.long().cpu()
    assert sorted(perm.tolist()) == list(range(E))
    rcv = ei[1].cpu()
    snd = ei[0].cpu()
    assert torch.equal(torch.bincount(snd, minlength=N), rowptr[1:] - rowptr[:-1])
    for n in range(N):
        seg = perm[rowptr[n] : rowptr[n + 1]]
        assert bool((rcv[seg] == n).all())


@pytest.mark.parametrize("name", NL_CASES)
def test_matches_reference_golden(cuda, name):
    from physicsml_b200.neighbour_list import construct_edge_indices_and_attrs, radius_graph

    g = torch.load(os.path.join(GOLD, name + ".pt"), weights_only=False)
    pos = g["pos"].to(cuda)
    cell = None if g["cell"] is None else g["cell"].to(cuda)
    ei, attrs, sh = construct_edge_indices_and_attrs(pos, g["cutoff"], None, None, g["pbc"], cell, False)
    assert attrs is None and ei.dtype == torch.int64 and sh.dtype == pos.dtype
    assert bool((ei[0][1:] >= ei[0][:-1]).all())  # sorted by sender like the reference (linked_cell.py:277-279)
    assert ei.shape == g["edge_index"].shape
    # our order IS (sender, receiver, shift), i.e. already the canonical order of the golden file: compare directly
    assert torch.equal(ei.cpu(), g["edge_index"].long())
    assert torch.equal(sh.cpu(), g["shift"].to(sh.dtype))
    _, _, plan = radius_graph(pos, g["cutoff"], pbc=g["pbc"], cell=cell)
    _check_plan(ei, plan, pos.shape[0])


@pytest.mark.parametrize("n,box,rc", [(3000, 31.07, 6.0), (700, 9.0, 6.0), (64, 5.0, 6.0)])
def test_pbc_against_oracle(cuda, n, box, rc):
    """uniform random atoms at water density; includes a cell smaller than the cutoff (multiple periodic images)"""
    from oracle.ref_nl import ref_radius_graph
    from physicsml_b200.neighbour_list import radius_graph

    g = torch.Generator().manual_seed(n)
    pos = torch.rand(n, 3, generator=g) * box
    cell = torch.eye(3) * box
    ei_ref, sh_ref = ref_radius_graph(pos, rc, (True, True, True), cell)
    ei, sh, plan = radius_graph(pos.to(cuda), rc, pbc=(True, True, True), cell=cell.to(cuda))
    assert ei.shape == ei_ref.shape
    assert torch.equal(ei.cpu(), ei_ref) and torch.equal(sh.cpu(), sh_ref)
    # symmetry (i, j, S) <-> (j, i, -S) is what the plan relies on
    assert int((plan.perm < 0).sum()) == 0


def test_open_boundary_batch(cuda):
    from oracle.ref_nl import ref_radius_graph_batched
    from physicsml_b200.neighbour_list import radius_graph

    pos, z, batch = helpers.random_molecules(12, 5, 30, 4, seed=7)
    ei_ref, _ = ref_radius_graph_batched(pos, batch, 5.0)
    ei, sh, plan = radius_graph(pos.to(cuda), 5.0, batch=batch.to(cuda))
    assert torch.equal(ei.cpu(), ei_ref)
    assert float(sh.abs().max()) == 0.0
    _check_plan(ei, plan, pos.shape[0])
    # same convention as helpers.radius_graph_n2 (what the model tests feed)
    assert torch.equal(ei.cpu(), helpers.radius_graph_n2(pos, batch, 5.0))


def test_known_answer_edge_counts(cuda):
    from physicsml_b200.neighbour_list import construct_edge_indices_and_attrs

    g = torch.load(os.path.join(GOLD, "gdb9_edges.pt"), weights_only=False)
    counts = [construct_edge_indices_and_attrs(p.to(cuda), 5.0, None, None, None, None, False)[0].shape[1] for p in g["positions"]]
    assert counts == g["edge_counts"]


def test_empty_and_isolated(cuda):
    from physicsml_b200.neighbour_list import radius_graph

    pos = torch.tensor([[0.0, 0, 0], [100.0, 0, 0]], device=cuda)
    ei, sh, plan = radius_graph(pos, 5.0)
    assert ei.shape == (2, 0) and sh.shape == (0, 3)
    ei, sh, plan = radius_graph(pos[:1], 5.0)
    assert ei.shape == (2, 0)


def test_periodic_model_parity(cuda):
    """MACE on a periodic cell, neighbour list + plan from the GPU kernels, vs the oracle on the reference-semantics graph"""
    from oracle.ref_model import RefPooledMACE, energy_and_forces
    from oracle.ref_nl import ref_radius_graph
    from physicsml_b200.mace import PooledMACEModule, compute_forces_by_gradient
    from physicsml_b200.neighbour_list import radius_graph

    kw = dict(cut_off=4.0, num_bessel=8, num_polynomial_cutoff=5, max_ell=3, num_interactions=2, num_node_feats=3,
              num_edge_feats=0, hidden_irreps="16x0e + 16x1o", avg_num_neighbours=20.0, correlation=3)
    torch.manual_seed(0)
    ref = RefPooledMACE(mlp_irreps="16x0e", **kw)
    mine = PooledMACEModule(mlp_irreps="16x0e", **kw)
    mine.load_state_dict(ref.state_dict())
    mine = mine.to(cuda)
    g = torch.Generator().manual_seed(3)
    n, box = 60, 8.5
    # jittered lattice so no two atoms are unphysically close
    grid = torch.stack(torch.meshgrid(*[torch.arange(4)] * 3, indexing="ij"), -1).reshape(-1, 3)[:n].float() * (box / 4)
    pos = grid + 0.3 * torch.randn(n, 3, generator=g) + 2.0  # partly outside the cell: exercises the wrap
    cell = torch.tensor([[box, 0, 0], [0.7, box, 0], [0.3, -0.4, box]])
    z = torch.randint(0, 3, (n,), generator=g)
    ei, sh = ref_radius_graph(pos, kw["cut_off"], (True, True, True), cell)
    data = {"coordinates": pos, "edge_index": ei, "cell": cell, "cell_shift_vector": sh,
            "node_attrs": torch.nn.functional.one_hot(z, 3).float(), "batch": torch.zeros(n, dtype=torch.long), "num_graphs": torch.tensor(1)}
    e_ref, f_ref = energy_and_forces(ref, data)
    d = helpers.to_device(data, cuda)
    ei_g, sh_g, plan = radius_graph(d["coordinates"], kw["cut_off"], pbc=(True, True, True), cell=d["cell"])
    assert torch.equal(ei_g.cpu(), ei) and torch.equal(sh_g.cpu(), sh)
    d["edge_index"], d["cell_shift_vector"], d["_edge_plan"] = ei_g, sh_g, plan
    d["coordinates"].requires_grad_(True)
    e = mine(d)["y_graph_scalars"]
    f = compute_forces_by_gradient(e, d["coordinates"])
    assert (e.cpu() - e_ref).abs().max() <= 1e-5 * e_ref.abs().max()
    assert (f.cpu() - f_ref).abs().max() <= 1e-4
    # and without the prebuilt plan (generic edge order path)
    d2 = helpers.to_device(data, cuda)
    d2["coordinates"].requires_grad_(True)
    e2 = mine(d2)["y_graph_scalars"]
    assert (e2.cpu() - e_ref).abs().max() <= 1e-5 * e_ref.abs().max()
