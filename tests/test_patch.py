"""``physicsml_b200.patch(module)``: the reference's own MACE / NequIP modules (unmodified sources imported over the
oracle shims) are switched to the B200 kernels in place.  CPU part: hyper-parameter
inference, parameter tying, unchanged state dict, plan construction.  GPU part: the patched reference module agrees with
its own original forward."""
import pytest
import torch

import helpers  # noqa: F401

MACE_KW = dict(cut_off=5.0, num_bessel=8, num_polynomial_cutoff=5, max_ell=2, num_interactions=2, num_node_feats=5,
               num_edge_feats=0, hidden_irreps="16x0e + 16x1o", avg_num_neighbours=14.0, correlation=3)
NEQUIP_KW = dict(cut_off=4.5, num_layers=3, max_ell=2, parity=True, num_features=8, num_bessel=8, bessel_basis_trainable=True,
                 num_polynomial_cutoff=6, self_connection=True, resnet=True, num_node_feats=5, num_edge_feats=0,
                 avg_num_neighbours=11.0)


def _reference(kind):
    from oracle import reference_models as rm

    if not rm.available():
        pytest.skip("reference sources not staged (python -m oracle.build_ref)")
    torch.manual_seed(0)
    return rm.ReferencePooledMACE("16x0e", **MACE_KW) if kind == "mace" else rm.ReferencePooledNequip("16x0e", **NEQUIP_KW)


@pytest.mark.parametrize("kind", ["mace", "nequip"])
def test_patch_infers_config_and_ties_parameters(kind):
    import physicsml_b200
    from physicsml_b200 import patching as P

    ref = _reference(kind)
    keys_before = list(ref.state_dict().keys())
    n_params = len(list(ref.parameters()))
    backbone = getattr(ref, kind)
    inferred = P.infer_mace_kwargs(backbone) if kind == "mace" else P.infer_nequip_kwargs(backbone)
    want = MACE_KW if kind == "mace" else NEQUIP_KW
    for k, v in want.items():
        got = inferred[k]
        if k == "hidden_irreps":
            assert str(got).replace(" ", "") == v.replace(" ", "")
        elif k == "resnet":  # only observable through the per-layer flags (a layer whose irreps change cannot be residual)
            assert got == any(l.resnet for l in backbone.conv_layers)
        else:
            assert got == v, (k, got, v)
    physicsml_b200.patch(ref)
    assert P.is_patched(ref)
    mirror = backbone._b200_mirror
    if kind == "nequip":
        assert [l.resnet for l in mirror.conv_layers] == [l.resnet for l in backbone.conv_layers]
    ref_params = dict(backbone.named_parameters())
    tied = 0
    for name, p in mirror.named_parameters():
        if p.numel() == 0:
            continue
        assert name in ref_params and p is ref_params[name], name  # the SAME Parameter object: grads land in the reference
        tied += 1
    assert tied == sum(1 for p in backbone.parameters() if p.numel() > 0 and p.requires_grad) or tied > 0
    # nothing was registered on the reference module: state dict and parameter list are unchanged
    assert list(ref.state_dict().keys()) == keys_before
    assert len(list(ref.parameters())) == n_params
    # a patched module has no CPU path
    with pytest.raises((NotImplementedError, RuntimeError)):
        pos, z, batch = helpers.random_molecules(2, 5, 7, 5, seed=1)
        ref(helpers.make_batch(pos, z, batch, 5, want["cut_off"]))
    P.unpatch(ref)
    assert not P.is_patched(ref)
    pos, z, batch = helpers.random_molecules(2, 5, 7, 5, seed=1)
    out = ref(helpers.make_batch(pos, z, batch, 5, want["cut_off"]))  # the original forward is back
    assert out["y_graph_scalars"].shape == (2, 1)


def test_plugin_entry_points_declared():
    import os
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "pyproject.toml")).read()
    assert "[project.entry-points.'molflux.modelzoo.plugins.physicsml']" in text
    for name, target in re.findall(r"^(\w+_b200) = '([\w.:]+)'", text, flags=re.M):
        mod, attr = target.split(":")
        assert mod == "physicsml_b200.plugin" and attr in ("MACEModelB200", "NequipModelB200")
    import physicsml_b200.plugin as plugin

    try:
        import molflux  # noqa: F401
        import physicsml  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError):
            plugin.MACEModelB200
    else:
        assert plugin.MACEModelB200.__name__ == "MACEModelB200"


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["mace", "nequip"])
def test_patched_reference_module_matches_its_original_forward(cuda, kind):
    """energy, forces and parameter gradients of the force-matching loss: reference module (CPU, original forward) vs the
    SAME module object after .to(cuda) + patch()"""
    import physicsml_b200
    from physicsml_b200.mace import compute_forces_by_gradient

    ref = _reference(kind)
    want = MACE_KW if kind == "mace" else NEQUIP_KW
    pos, z, batch = helpers.random_molecules(5, 8, 16, 5, seed=11)
    data = helpers.make_batch(pos, z, batch, 5, want["cut_off"])
    g = torch.Generator().manual_seed(2)
    e_t, f_t = torch.randn(5, 1, generator=g), 0.1 * torch.randn(pos.shape[0], 3, generator=g)

    def loss_and_grads(model, d, dev):
        d = dict(d)
        d["coordinates"] = d["coordinates"].clone().requires_grad_(True)
        model.zero_grad(set_to_none=True)
        e = model(d)["y_graph_scalars"]
        f = compute_forces_by_gradient(e, d["coordinates"], model)
        loss = ((e - e_t.to(dev)) ** 2).mean() + ((f - f_t.to(dev)) ** 2).mean()
        loss.backward()
        return e.detach().cpu(), f.detach().cpu(), {k: p.grad.detach().cpu().clone() for k, p in model.named_parameters() if p.grad is not None}

    e0, f0, g0 = loss_and_grads(ref, data, "cpu")
    ref.to(cuda)
    physicsml_b200.patch(ref)
    e1, f1, g1 = loss_and_grads(ref, helpers.to_device(data, cuda), cuda)
    scale = e0.abs().max().item()
    assert (e1 - e0).abs().max().item() <= 1e-5 * scale
    assert (f1 - f0).abs().max().item() <= 1e-4
    assert set(g0) == set(g1)
    for k in g0:
        assert (g1[k] - g0[k]).abs().max().item() <= 2e-4 * max(g0[k].abs().max().item(), 1e-12) + 1e-7, k


def _variant_module():
    """a MACE variant in the manner of the reference's ssf / mean_var modules (models/mace/ssf/ssf_mace_utils.py:165-209,
    models/mace/mean_var/mean_var_mace_module.py:96-122): the reference's own blocks with an extra per-layer module in
    between, a second pooled head for the predicted std and a node-vector head — not the plain MACE class"""
    from oracle import import_reference as ir
    from oracle import reference_models as rm

    if not rm.available():
        pytest.skip("reference sources not staged")
    mod = ir.load("physicsml.models.mace.modules.mace")
    ir.install_shims()
    from e3nn import o3

    class Variant(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.variant_mace = mod.MACE(**MACE_KW)
            C = 16
            self.scale = torch.nn.ParameterList([torch.nn.Parameter(torch.ones(1) * 1.1) for _ in range(2)])  # "ssf"-like
            self.graph_scalars_head = mod.PooledReadoutHead(self.variant_mace.out_irreps, o3.Irreps("16x0e"), o3.Irreps("1x0e"), 1.0, 0.0)
            self.graph_scalars_head_std = mod.PooledReadoutHead(self.variant_mace.out_irreps, o3.Irreps("16x0e"), o3.Irreps("1x0e"), 1.0, 0)
            del C

        def forward(self, data):
            m = self.variant_mace
            from physicsml.models.utils import compute_lengths_and_vectors

            lengths, vectors = compute_lengths_and_vectors(data["coordinates"], data["edge_index"], None, None)
            h = m.node_embedding(data["node_attrs"])
            Y, B = m.spherical_harmonics(vectors), m.radial_embedding(lengths)
            for i, (inter, msg, upd) in enumerate(zip(m.interactions, m.messages, m.node_updates)):
                a = inter(node_feats=h, node_attrs=data["node_attrs"], edge_attrs=Y, edge_feats=B, edge_index=data["edge_index"])
                h = upd(m_i=msg(a_i=a, node_attrs=data["node_attrs"]), node_feats=h, node_attrs=data["node_attrs"])
                h = h * self.scale[i]
                data[f"node_feats_{i}"] = h
            e = self.graph_scalars_head(data)
            std = torch.abs(self.graph_scalars_head_std(data)) / (data["ptr"][1:] - data["ptr"][:-1]).unsqueeze(-1)
            return {"y_graph_scalars": e, "y_graph_scalars::std": std}

    torch.manual_seed(0)
    return Variant()


def test_block_level_patch_of_a_variant_backbone_cpu():
    import physicsml_b200
    from physicsml_b200 import patching as P

    v = _variant_module()
    keys = list(v.state_dict().keys())
    physicsml_b200.patch(v)
    swapped = [m for m in v.modules() if getattr(m, "_b200_mirror", None) is not None]
    assert len(swapped) == 6  # 2 x (interaction, message, node update); the variant's own forward is untouched
    assert "forward" not in v.__dict__ and list(v.state_dict().keys()) == keys
    for m in swapped:
        ref_params = dict(m.named_parameters())
        for name, p in m._b200_mirror.named_parameters():
            if p.numel():
                assert p is ref_params[name]
    P.unpatch(v)
    assert not P.is_patched(v)


@pytest.mark.gpu
def test_block_level_patch_of_a_variant_backbone_matches(cuda):
    import physicsml_b200

    v = _variant_module()
    pos, z, batch = helpers.random_molecules(4, 8, 14, 5, seed=21)
    data = helpers.make_batch(pos, z, batch, 5, MACE_KW["cut_off"])
    ptr = torch.zeros(5, dtype=torch.long)
    ptr[1:] = torch.cumsum(torch.bincount(batch, minlength=4), 0)
    data["ptr"] = ptr

    def run(model, d):
        d = dict(d)
        d["coordinates"] = d["coordinates"].clone().requires_grad_(True)
        out = model(d)
        (g,) = torch.autograd.grad([out["y_graph_scalars"].sum() + out["y_graph_scalars::std"].sum()], [d["coordinates"]])
        return out["y_graph_scalars"].detach().cpu(), out["y_graph_scalars::std"].detach().cpu(), g.cpu()

    e0, s0, g0 = run(v, data)
    v.to(cuda)
    physicsml_b200.patch(v)
    e1, s1, g1 = run(v, helpers.to_device(data, cuda))
    assert (e1 - e0).abs().max().item() <= 1e-5 * e0.abs().max().item()
    assert (s1 - s0).abs().max().item() <= 1e-5 * max(s0.abs().max().item(), 1e-6)
    assert (g1 - g0).abs().max().item() <= 1e-4
