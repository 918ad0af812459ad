"""GPU neighbour list (K1, through the C ABI) against the oracle restatement of the reference's semantics and the
golden edge sets produced by the reference itself.  Bar: the edge set is bit-exact after sorting."""
import glob
import os

import pytest
from typing import Optional
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


def test_bond_feature_merge_matches_reference_golden(cuda):
    """neighbourhood_list_torch.py:98-156: supplied bond edges + attributes merged into the radius graph;
    golden produced by the reference's own construct_edge_indices_and_attrs (tests/golden/make_golden.py §5)"""
    from physicsml_b200.neighbour_list import construct_edge_indices_and_attrs

    g = torch.load(os.path.join(GOLD, "bondmerge_nl.pt"), weights_only=False)
    ei, ea, sh = construct_edge_indices_and_attrs(g["pos"].to(cuda), g["cutoff"], g["bonds"].to(cuda), g["attrs"].to(cuda), None, None, False)
    assert torch.equal(ei.cpu(), g["open"]["edge_index"]) and torch.equal(ea.cpu(), g["open"]["edge_attrs"])
    assert torch.equal(sh.cpu(), g["open"]["shift"])
    p = g["pbc"]
    ei, ea, sh = construct_edge_indices_and_attrs((g["pos"] + p["offset"]).to(cuda), g["cutoff"], g["bonds"].to(cuda), g["attrs"].to(cuda),
                                                  (True, True, True), p["cell"].to(cuda), False)
    assert torch.equal(ei.cpu(), p["edge_index"]) and torch.equal(ea.cpu(), p["edge_attrs"]) and torch.equal(sh.cpu(), p["shift"])
    with pytest.raises(ValueError):  # a bond longer than the cut-off is not a subset of the radius graph
        construct_edge_indices_and_attrs(g["pos"].to(cuda), 1.0, g["bonds"].to(cuda), g["attrs"].to(cuda), None, None, False)


def test_batched_periodic_neighbour_list(cuda):
    """a collated batch of periodic structures in ONE call == one call per structure, for a shared cell
    (the reference's batch contract) and for one (triclinic) cell per structure with cartesian shifts"""
    from oracle.ref_nl import ref_radius_graph, sorted_edge_set
    from physicsml_b200.neighbour_list import collate, radius_graph

    gen = torch.Generator().manual_seed(12)
    rc = 4.0
    cells = [torch.eye(3) * 9.0, torch.tensor([[8.0, 0, 0], [2.0, 7.5, 0], [-1.0, 1.0, 10.0]]), torch.eye(3) * 5.5]
    ns = [30, 41, 12]
    structs = []
    for c, n in zip(cells, ns):
        frac = torch.rand(n, 3, generator=gen) * 1.4 - 0.2  # some atoms outside the cell: exercises the wrap + shift fix-up
        structs.append({"coordinates": frac @ c, "species": torch.randint(0, 3, (n,), generator=gen), "cell": c})
    # (a) per-structure cells
    data = collate(structs, rc, cuda, num_species=3)
    off = 0
    for s in structs:
        n = s["coordinates"].shape[0]
        ei_ref, sh_ref = ref_radius_graph(s["coordinates"], rc, (True, True, True), s["cell"])
        m = (data["edge_index"][0] >= off) & (data["edge_index"][0] < off + n)
        ei = (data["edge_index"][:, m] - off).cpu()
        assert bool(((ei >= 0) & (ei < n)).all())  # no edge leaves its structure
        cart_ref = sh_ref @ s["cell"]
        a = sorted_edge_set(ei_ref, sh_ref)[0]
        # same edges in the same (sender, receiver, shift) order; shifts compared in length units
        assert ei.shape == ei_ref.shape
        key = lambda e, c: sorted(zip(e[0].tolist(), e[1].tolist(), [tuple(round(x, 3) for x in r) for r in c.tolist()]))  # noqa: E731
        assert key(ei, data["cell_shift_vector"][m].cpu()) == key(ei_ref, cart_ref)
        off += n
    assert torch.equal(data["cell"].cpu(), torch.eye(3))
    # (b) one shared cell: integer shifts, identical to single-structure calls
    shared = [dict(s, cell=cells[1]) for s in structs]
    data = collate(shared, rc, cuda, num_species=3)
    assert torch.equal(data["cell"].cpu(), cells[1])
    off = 0
    for s in shared:
        n = s["coordinates"].shape[0]
        ei1, sh1, _ = radius_graph(s["coordinates"].to(cuda), rc, pbc=(True, True, True), cell=cells[1].to(cuda))
        m = (data["edge_index"][0] >= off) & (data["edge_index"][0] < off + n)
        assert torch.equal(data["edge_index"][:, m] - off, ei1) and torch.equal(data["cell_shift_vector"][m], sh1)
        off += n
    _check_plan(data["edge_index"], data["_edge_plan"], off)


def test_collated_batch_runs_the_model(cuda):
    """collate() output is the reference's batch dict: the model consumes it directly (periodic, per-structure cells)"""
    from oracle.ref_model import RefPooledMACE, energy_and_forces
    from physicsml_b200 import workloads as wl
    from physicsml_b200.mace import PooledMACEModule
    from physicsml_b200.neighbour_list import collate

    kw = dict(cut_off=4.0, num_bessel=8, num_polynomial_cutoff=5, max_ell=2, num_interactions=2, num_node_feats=3, num_edge_feats=0,
              hidden_irreps="16x0e + 16x1o", avg_num_neighbours=20.0, correlation=3)
    torch.manual_seed(0)
    ref = RefPooledMACE(mlp_irreps="16x0e", **kw)
    mine = PooledMACEModule(mlp_irreps="16x0e", **kw)
    mine.load_state_dict(ref.state_dict(), strict=True)
    mine = mine.to(cuda)
    gen = torch.Generator().manual_seed(3)
    structs = []
    for a, n in ((7.0, 20), (8.5, 28)):
        c = torch.eye(3) * a
        structs.append({"coordinates": torch.rand(n, 3, generator=gen) * a, "species": torch.randint(0, 3, (n,), generator=gen), "cell": c})
    data = collate(structs, 4.0, cuda, num_species=3)
    e, f = wl.energy_forces(mine, data)
    cpu = {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in data.items() if k != "_edge_plan"}
    e_ref, f_ref = energy_and_forces(ref, cpu)
    scale = e_ref.abs().max().item()
    assert (e.cpu() - e_ref).abs().max().item() <= 1e-5 * scale and (f.cpu() - f_ref).abs().max().item() <= 1e-4


def test_md_loop_skin_list_is_exact_and_reused(cuda):
    """MDCalculator (Verlet skin + optional CUDA-graph replay) over a short random walk of a periodic
    box == a fresh neighbour list at the model cut-off + eager evaluation at every step; the list is rebuilt only when
    an atom moved more than skin / 2"""
    from physicsml_b200 import workloads as wl
    from physicsml_b200.mace import PooledMACEModule
    from physicsml_b200.md import MDCalculator
    from physicsml_b200.neighbour_list import radius_graph

    kw = dict(cut_off=4.0, num_bessel=8, num_polynomial_cutoff=5, max_ell=2, num_interactions=2, num_node_feats=3, num_edge_feats=0,
              hidden_irreps="16x0e + 16x1o", avg_num_neighbours=20.0, correlation=3)
    torch.manual_seed(0)
    model = PooledMACEModule(mlp_irreps="16x0e", **kw).to(cuda)
    gen = torch.Generator().manual_seed(8)
    N, a = 96, 11.0
    cell = torch.eye(3) * a
    pos = torch.rand(N, 3, generator=gen) * a
    attrs = torch.nn.functional.one_hot(torch.randint(0, 3, (N,), generator=gen), 3).float()
    for use_graph in (False, True):
        md = MDCalculator(model, 4.0, attrs, cell=cell, skin=0.6, use_graph=use_graph)
        p = pos.clone()
        for step in range(12):
            e, f = md.energy_forces(p)
            ei, sh, plan = radius_graph(p.to(cuda), 4.0, pbc=(True, True, True), cell=cell.to(cuda))
            d = {"coordinates": p.to(cuda), "edge_index": ei, "cell": cell.to(cuda), "cell_shift_vector": sh, "node_attrs": attrs.to(cuda),
                 "batch": torch.zeros(N, dtype=torch.long, device=cuda), "num_graphs": torch.tensor(1)}
            e0, f0 = wl.energy_forces(model, d)
            assert (e - e0.reshape(-1)).abs().max().item() <= 2e-6 * max(e0.abs().max().item(), 1e-3), (use_graph, step)
            assert (f - f0).abs().max().item() <= 2e-6
            assert md.n_edges >= ei.shape[1]
            p = p + 0.04 * torch.randn(N, 3, generator=gen)  # ~0.07 A per step: a rebuild every few steps
        assert 2 <= md.n_builds < 12, md.n_builds
    # eager default (large-system mode): no skin, the exact list is rebuilt on the GPU at every step
    md = MDCalculator(model, 4.0, attrs, cell=cell, use_graph=False)
    assert md.skin == 0.0
    for step in range(3):
        md.energy_forces(pos + 0.01 * step)
    assert md.n_builds == 3
    # open boundaries + unit scaling as the OpenMM module applies it (nm in, kJ/mol out)
    md = MDCalculator(model, 4.0, attrs, skin=0.5, position_scaling=10.0, output_scaling=96.4853)
    e, f = md.energy_forces(pos / 10.0)
    ei, sh, plan = radius_graph(pos.to(cuda), 4.0)
    d = {"coordinates": pos.to(cuda), "edge_index": ei, "node_attrs": attrs.to(cuda), "batch": torch.zeros(N, dtype=torch.long, device=cuda),
         "num_graphs": torch.tensor(1)}
    e0, f0 = wl.energy_forces(model, d)
    assert (e - 96.4853 * e0.reshape(-1)).abs().max().item() <= 1e-5 * 96.4853 * max(e0.abs().max().item(), 1e-3)
    assert (f - 964.853 * f0).abs().max().item() <= 1e-3
    r = md.ase_results((pos / 10.0).numpy())
    assert isinstance(r["energy"], float) and r["forces"].shape == (N, 3)


def test_torchscript_neighbour_list_op(cuda):
    """the TORCH_LIBRARY op, called from a scripted function, returns the same edges as the Python entry point"""
    from physicsml_b200 import build
    from physicsml_b200.neighbour_list import radius_graph

    torch.ops.load_library(build.TORCH_LIB)

    @torch.jit.script
    def edges(pos: torch.Tensor, cell: Optional[torch.Tensor], cutoff: float):
        return torch.ops.physicsml_b200_ts.radius_graph(pos, cell, cutoff, False)

    gen = torch.Generator().manual_seed(4)
    cell = torch.tensor([[9.0, 0, 0], [1.0, 8.0, 0], [0.5, -1.0, 10.0]])
    pos = (torch.rand(150, 3, generator=gen) @ cell).to(cuda)
    ei, sh = edges(pos, cell.to(cuda), 4.5)
    ei0, sh0, _ = radius_graph(pos, 4.5, pbc=(True, True, True), cell=cell.to(cuda))
    assert torch.equal(ei, ei0) and torch.equal(sh, sh0)
    ei, sh = edges(pos, None, 4.5)
    ei0, sh0, _ = radius_graph(pos, 4.5)
    assert torch.equal(ei, ei0) and float(sh.abs().max()) == 0.0
