"""Host-side logic that needs no GPU: irreps / path tables / plans against the oracle's e3nn restatement, state-dict
compatibility with the reference layout, generated-kernel registry coverage, loud failure without CUDA."""
import numpy as np
import pytest
import torch

import helpers  # noqa: F401


def test_tp_path_table_matches_oracle_instructions():
    from e3nn import o3
    from oracle.ref_model import uvu_instructions
    from physicsml_b200 import so3

    for node, lmax in [("4x0e", 2), ("4x0e+4x1o", 2), ("128x0e+128x1o", 3), ("256x0e+256x1o+256x2e", 3)]:
        C = int(node.split("x")[0])
        sh = o3.Irreps.spherical_harmonics(lmax)
        target = (sh * C).sort().irreps.simplify()
        irr, ins = uvu_instructions(o3.Irreps(node), sh, target)
        spec = so3.build_tp_spec(node, lmax, str(target))
        assert spec.n_paths == len(ins)
        assert [(q.i_in, q.l2, q.i_out) for q in spec.paths] == [(i, j, k) for i, j, k, _, _ in ins]
        assert spec.d_out * C == irr.dim and spec.d_in * C == o3.Irreps(node).dim


def test_tp_sparse_terms_equal_dense_oracle():
    """the generated code's term list evaluated in numpy == oracle o3.TensorProduct (dense einsum over w3j)"""
    from e3nn import o3
    from oracle.ref_model import uvu_instructions
    from physicsml_b200 import so3

    node, lmax, C = "2x0e+2x1o", 3, 2
    sh = o3.Irreps.spherical_harmonics(lmax)
    target = (sh * C).sort().irreps.simplify()
    irr, ins = uvu_instructions(o3.Irreps(node), sh, target)
    tp = o3.TensorProduct(o3.Irreps(node), sh, irr, instructions=ins, shared_weights=False, internal_weights=False)
    spec = so3.build_tp_spec(node, lmax, str(target))
    g = torch.Generator().manual_seed(0)
    h = torch.randn(1, C * spec.d_in, generator=g, dtype=torch.float64)
    Y = torch.randn(1, spec.n_sh, generator=g, dtype=torch.float64)
    R = torch.randn(1, spec.n_paths * C, generator=g, dtype=torch.float64)
    ref = tp(h, Y, R)[0].numpy()
    out = np.zeros(C * spec.d_out)
    in_l = [l for l, _ in spec.node_ls]
    out_l = [l for l, _ in spec.out_ls]
    in_blk_off = np.concatenate([[0], np.cumsum([2 * l + 1 for l in in_l])])
    out_blk_off = np.concatenate([[0], np.cumsum([2 * l + 1 for l in out_l])])
    for p, o, i, j, c in so3.tp_terms(spec):
        bi = np.searchsorted(in_blk_off, i, side="right") - 1
        bo = np.searchsorted(out_blk_off, o, side="right") - 1
        di, do = 2 * in_l[bi] + 1, 2 * out_l[bo] + 1
        for u in range(C):
            hv = h[0, C * in_blk_off[bi] + u * di + (i - in_blk_off[bi])].item()
            out[C * out_blk_off[bo] + u * do + (o - out_blk_off[bo])] += R[0, p * C + u].item() * c * hv * Y[0, j].item()
    assert np.abs(out - ref).max() < 1e-12


def test_contraction_monomials_equal_dense_oracle():
    from e3nn import o3
    from oracle.ref_model import _SymContraction
    from physicsml_b200 import so3

    lmax, C, Z, corr = 2, 2, 3, 3
    inter = (o3.Irreps.spherical_harmonics(lmax) * C).sort().irreps.simplify()
    torch.manual_seed(0)
    ref = _SymContraction(inter, o3.Irreps("2x0e+2x1o"), corr, Z).double()
    S = (lmax + 1) ** 2
    x = torch.randn(1, C, S, dtype=torch.float64)
    y = torch.nn.functional.one_hot(torch.tensor([1]), Z).double()
    want = ref(x, y)[0].detach().numpy()
    coup = tuple((l, (-1) ** l) for l in range(lmax + 1))
    got, off = np.zeros_like(want), 0
    for a, out in enumerate([(0, 1), (1, -1)]):
        con = ref.contractions[a]
        ws = {corr: con.weights_max}
        for i, w in enumerate(con.weights):
            ws[corr - 1 - i] = w
        d = 2 * out[0] + 1
        for nu in range(1, corr + 1):
            for (k, M, idx), c in so3.contraction_monomials(coup, out, nu).items():
                for ch in range(C):
                    got[off + ch * d + M] += ws[nu][1, k, ch].item() * c * np.prod([x[0, ch, i].item() for i in idx])
        off += C * d
    assert np.abs(got - want).max() < 1e-6 * max(1.0, np.abs(want).max())  # U is built in fp32 by the reference


def test_reduced_contraction_basis_is_exact():
    """K6 evaluates W' = B^T W on a basis of the (linearly dependent) per-weight polynomials: the reduced form must
    reproduce the full sum  sum_k W_k P_k[M](x)  for arbitrary weights, and the registry must carry the reduced sizes."""
    from physicsml_b200 import _lib, so3
    from physicsml_b200.codegen import SymSpec

    rng = np.random.default_rng(5)
    lmax = 3
    coup = tuple((l, (-1) ** l) for l in range(lmax + 1))
    x = rng.standard_normal((lmax + 1) ** 2)
    total_full = total_red = 0
    for out in [(0, 1), (1, -1), (2, 1)]:
        for nu in (1, 2, 3):
            full = so3.contraction_monomials(coup, out, nu)
            red, B, sel = so3.reduced_contraction(coup, out, nu)
            K = so3.u_matrix(coup, out, nu).shape[-1]
            assert B.shape == (K, len(sel)) and len(sel) <= K
            W = rng.standard_normal(K)
            Wr = B.T @ W
            want, got = np.zeros(2 * out[0] + 1), np.zeros(2 * out[0] + 1)
            for (k, M, idx), c in full.items():
                want[M] += W[k] * c * np.prod(x[list(idx)])
            for (j, M, idx), c in red.items():
                got[M] += Wr[j] * c * np.prod(x[list(idx)])
            assert np.abs(got - want).max() < 1e-12 * max(1.0, np.abs(want).max())
            total_full += len(full)
            total_red += len(red)
    assert total_red * 3 < total_full  # 5911 -> 1211 monomial terms at lmax 3, L <= 2
    spec = SymSpec(3, [(0, 1), (1, -1)], 3)
    assert spec.K_full == [[1, 4, 23], [1, 6, 51]] and spec.K == [[1, 4, 8], [1, 3, 12]]
    assert _lib.registry()["sym"][spec.key]["K"] == spec.K


def test_generated_registry_covers_baseline_configs():
    from physicsml_b200 import _lib, so3

    reg = _lib.registry()
    for node, lmax in [("32x0e", 2), ("32x0e+32x1o", 2), ("128x0e", 3), ("128x0e+128x1o", 3), ("256x0e", 3), ("256x0e+256x1o+256x2e", 3)]:
        tgt = so3.irreps_str([(1, l, (-1) ** l) for l in range(lmax + 1)])
        assert so3.build_tp_spec(node, lmax, tgt).key in reg["tp"]
    for key in ("l2__o0e_1o__c3", "l2__o0e__c3", "l3__o0e_1o__c3", "l3__o0e__c3", "l3__o0e_1o_2e__c3", "l2__o0e_1o__c2"):
        assert key in reg["sym"]


def test_product_state_dict_is_reference_layout():
    from oracle.ref_model import RefPooledMACE
    from physicsml_b200.mace import PooledMACEModule

    kw = dict(cut_off=5.0, num_bessel=8, num_polynomial_cutoff=5, max_ell=3, num_interactions=2, num_node_feats=10,
              num_edge_feats=0, hidden_irreps="16x0e + 16x1o", avg_num_neighbours=10.0, correlation=3)
    a, b = PooledMACEModule(mlp_irreps="16x0e", **kw), RefPooledMACE(mlp_irreps="16x0e", **kw)
    sa, sb = a.state_dict(), b.state_dict()
    assert set(sa) == set(sb)
    assert all(sa[k].shape == sb[k].shape for k in sa)
    a.load_state_dict(sb, strict=True)
    for k in sa:
        if "U_matrices" in k:
            assert (sa[k] - sb[k]).abs().max() < 1e-6


def test_linear_plan_matches_oracle_layout():
    from e3nn import o3
    from physicsml_b200.plans import ElementLinearPlan, LinearPlan

    for a, b in [("64x0e+96x1o+64x2e", "32x0e+32x1o+32x2e"), ("32x0e+32x1o", "1x0e"), ("4x0e", "4x0e")]:
        assert LinearPlan(a, b).weight_numel == o3.Linear(o3.Irreps(a), o3.Irreps(b)).weight_numel
    ref = o3.FullyConnectedTensorProduct(o3.Irreps("4x0e+4x1o+4x2e"), o3.Irreps("4x0e"), o3.Irreps("4x0e+4x1o+4x2e"))
    assert ElementLinearPlan("4x0e+4x1o+4x2e", 4, "4x0e+4x1o+4x2e").weight_numel == ref.weight_numel == 192


def _run_gemm_table(probs, A, B, C, M_runtime):
    """plain-numpy executor of a pml_gemm_problem table (the semantics both GEMM kernels implement)"""
    for p in probs:
        M = M_runtime if p.M < 0 else p.M
        for z in range(p.nbatch):
            acc = np.zeros((M, p.N))
            for c in range(p.nchunks):
                K = M_runtime if p.kc[c] < 0 else p.kc[c]
                a_idx = p.a_off[c] + z * p.a_bs + np.arange(M)[:, None] * p.a_rs + np.arange(K)[None, :] * p.a_cs
                b_idx = p.b_off[c] + z * p.b_bs + np.arange(K)[:, None] * p.b_rs + np.arange(p.N)[None, :] * p.b_cs
                acc += p.cscale[c] * (A[a_idx] @ B[b_idx])
            c_idx = p.c_off + z * p.c_bs + np.arange(M)[:, None] * p.c_rs + np.arange(p.N)[None, :] * p.c_cs
            if p.accumulate or (p.c_bs == 0 and p.nbatch > 1):
                np.add.at(C, c_idx, p.alpha * acc)
            else:
                C[c_idx] = p.alpha * acc


def test_linear_plan_component_major_tables_address_the_same_elements():
    """LinearPlan(in_layout="component") (the layout K4 writes for MACE's per-irrep linear): forward, dx and dW tables
    executed in numpy on a component-major input equal the irreps-layout tables on the e3nn-ordered input"""
    from physicsml_b200 import so3
    from physicsml_b200.plans import LinearPlan

    rng = np.random.default_rng(9)
    irr_in, irr_out, n = "6x0e+6x0e+6x1o+6x1o+6x1o+6x2e", "6x0e+6x1o+6x2e", 5
    pe = LinearPlan(irr_in, irr_out, scale=0.7)
    pc = LinearPlan(irr_in, irr_out, scale=0.7, in_layout="component")
    assert pe.weight_numel == pc.weight_numel
    x = rng.standard_normal((n, pe.d_in))
    xc, off = np.empty_like(x), 0
    for m, l, _ in so3.parse_irreps(irr_in):  # [mul, 2l+1] -> [2l+1, mul] inside every block
        d = 2 * l + 1
        xc[:, off:off + m * d] = x[:, off:off + m * d].reshape(n, m, d).transpose(0, 2, 1).reshape(n, m * d)
        off += m * d
    w = rng.standard_normal(pe.weight_numel)
    g = rng.standard_normal((n, pe.d_out))
    ye, yc = np.zeros(n * pe.d_out), np.zeros(n * pe.d_out)
    _run_gemm_table(pe._fwd, x.ravel(), w, ye, n)
    _run_gemm_table(pc._fwd, xc.ravel(), w, yc, n)
    assert np.abs(ye).max() > 0 and np.allclose(ye, yc, rtol=0, atol=1e-12)
    dxe, dxc = np.zeros(n * pe.d_in), np.zeros(n * pe.d_in)
    _run_gemm_table(pe._bwd_x, g.ravel(), w, dxe, n)
    _run_gemm_table(pc._bwd_x, g.ravel(), w, dxc, n)
    back, off = np.empty((n, pe.d_in)), 0
    dxc = dxc.reshape(n, -1)
    for m, l, _ in so3.parse_irreps(irr_in):  # component-major cotangent -> e3nn order
        d = 2 * l + 1
        back[:, off:off + m * d] = dxc[:, off:off + m * d].reshape(n, d, m).transpose(0, 2, 1).reshape(n, m * d)
        off += m * d
    assert np.allclose(dxe.reshape(n, -1), back, rtol=0, atol=1e-12)
    dwe, dwc = np.zeros(pe.weight_numel), np.zeros(pe.weight_numel)
    _run_gemm_table(pe._bwd_w, x.ravel(), g.ravel(), dwe, n)
    _run_gemm_table(pc._bwd_w, xc.ravel(), g.ravel(), dwc, n)
    assert np.abs(dwe).max() > 0 and np.allclose(dwe, dwc, rtol=0, atol=1e-12)


def test_element_linear_segment_tables_equal_the_dense_definition():
    """ElementLinearPlan's template problems, placed on the element groups the way pml_segment_problems does it
    (M / kc = group size, offsets + group start x row stride) and executed in numpy on the gathered component-major copies,
    equal out[n, w, m] = alpha sum_u W[u, z_n, w] x[n, u, m], its dx and its dW - incl. an element that does not occur"""
    import copy

    from physicsml_b200 import so3
    from physicsml_b200.plans import ElementLinearPlan

    rng = np.random.default_rng(3)
    irr_in, irr_out, Z, n = "4x0e+4x1o+4x1e", "4x0e+4x1e+4x1o+4x2e", 4, 23
    pl = ElementLinearPlan(irr_in, Z, irr_out)
    z = rng.integers(0, Z, n)
    z[z == 2] = 0  # element 2 is absent: an empty group
    order = np.argsort(z, kind="stable")
    seg = np.concatenate([[0], np.cumsum(np.bincount(z, minlength=Z))])
    x = rng.standard_normal((n, pl.d_in))
    g = rng.standard_normal((n, pl.d_out))
    w = rng.standard_normal(pl.weight_numel)

    def place(tmpl, meta):
        out = []
        for q0, m in zip(tmpl, meta):
            q = copy.copy(q0)
            start, count = int(seg[m.v]), int(seg[m.v + 1] - seg[m.v])
            q.a_off[0] += start * m.a_mult
            q.b_off[0] += start * m.b_mult
            q.c_off += start * m.c_mult
            if m.mode == 0:
                q.M = count
            else:
                q.kc[0] = count
            out.append(q)
        return out

    N = pl.seg_nprob
    fwd, bx, bw = (place(pl._seg_tmpl[k * N:(k + 1) * N], pl._seg_meta[k * N:(k + 1) * N]) for k in range(3))
    xs, gs = x[order][:, pl._cm_in], g[order][:, pl._cm_out]  # rows grouped by element, blocks component-major
    # dense definition
    want, want_dx, want_dw = np.zeros((n, pl.d_out)), np.zeros((n, pl.d_in)), np.zeros(pl.weight_numel)
    in_off, out_off = so3.irreps_offsets(pl.irreps_in), so3.irreps_offsets(pl.irreps_out)
    P = pl.c_plan
    for k in range(P.npaths):
        d, mi, mo = P.d[k], P.mul_in[k], P.mul_out[k]
        W = w[P.w_off[k]:P.w_off[k] + mi * Z * mo].reshape(mi, Z, mo)
        dW = want_dw[P.w_off[k]:P.w_off[k] + mi * Z * mo].reshape(mi, Z, mo)
        for b in range(n):
            xb = x[b, P.x_off[k]:P.x_off[k] + mi * d].reshape(mi, d)
            gb = g[b, P.o_off[k]:P.o_off[k] + mo * d].reshape(mo, d)
            want[b, P.o_off[k]:P.o_off[k] + mo * d] += (P.alpha[k] * np.einsum("uw,um->wm", W[:, z[b]], xb)).ravel()
            want_dx[b, P.x_off[k]:P.x_off[k] + mi * d] += (P.alpha[k] * np.einsum("uw,wm->um", W[:, z[b]], gb)).ravel()
            dW[:, z[b]] += P.alpha[k] * np.einsum("um,wm->uw", xb, gb)
    ys = np.zeros(n * pl.d_out_c)
    _run_gemm_table(fwd, xs.ravel(), w, ys, n)
    out = np.zeros((n, pl.d_out))
    out[np.ix_(order, pl._cm_out)] = ys.reshape(n, pl.d_out_c)
    assert np.abs(want).max() > 0 and np.allclose(out, want, rtol=0, atol=1e-12)
    dxs = np.zeros(n * pl.d_in_c)
    _run_gemm_table(bx, gs.ravel(), w, dxs, n)
    dx = np.zeros((n, pl.d_in))
    dx[np.ix_(order, pl._cm_in)] = dxs.reshape(n, pl.d_in_c)
    assert np.allclose(dx, want_dx, rtol=0, atol=1e-12)
    dw = np.zeros(pl.weight_numel)
    _run_gemm_table(bw, xs.ravel(), gs.ravel(), dw, n)
    assert np.abs(want_dw).max() > 0 and np.allclose(dw, want_dw, rtol=0, atol=1e-12)


def test_gemm_scheduling_helpers_and_grad_pruning_flag():
    """host-side scheduling of the grouped GEMM (whole waves, no nearly empty extra wave) and the context manager that
    switches weight gradients off for the force pass"""
    from physicsml_b200 import ops

    assert ops._tc_slots(128, False) == 296 and ops._tc_slots(64, False) == 444 and ops._tc_slots(64, True) == 296
    # dW of the wide radial layer: 10 column tiles on two waves of 296 slots -> 59 splits (590 CTAs), not 60 (600)
    assert ops._splitk_for(41570, 10, 2 * ops._tc_slots(128, False)) == 59
    assert ops._splitk_for(41570, 4, ops._tc_slots(64, False)) == 111
    assert ops._splitk_for(100, 10, 592) == 1  # never fewer than 64 k per CTA
    sk = ops._splitk_rows(325, 1280, 296)  # 325 row tiles on 296 slots: split the long reduction
    assert 2 <= sk <= 8 and 1280 // sk >= 128
    assert ops._splitk_rows(3000, 1280, 296) == 1 and ops._splitk_rows(325, 128, 296) == 1
    assert ops._PARAM_GRADS[0] is True
    with pytest.raises(RuntimeError):
        with ops.input_grads_only():
            assert ops._PARAM_GRADS[0] is False
            with ops.input_grads_only():
                assert ops._PARAM_GRADS[0] is False
            assert ops._PARAM_GRADS[0] is False
            raise RuntimeError("boom")
    assert ops._PARAM_GRADS[0] is True


def test_no_cpu_fallback():
    from physicsml_b200 import neighbour_list, ops

    with pytest.raises((NotImplementedError, RuntimeError)):
        ops.act(torch.randn(4), 0, 1.0)
    with pytest.raises(RuntimeError):
        neighbour_list.radius_graph(torch.randn(4, 3), 5.0)


def test_nequip_mirror_loads_reference_checkpoint_strict():
    """every parameter / buffer name and shape of the reference's NequIP fixture checkpoint
    (tests/data/nequip_model/...ckpt, copied into tests/golden/nequip_fixture_ckpt.pt) exists in the mirror"""
    import os

    from physicsml_b200.nequip import PooledNequipModule

    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "nequip_fixture_ckpt.pt"), weights_only=False)
    model = PooledNequipModule(mlp_irreps=g["mlp"], **g["kw"])
    res = model.load_state_dict(g["state_dict"], strict=False)
    assert not res.unexpected_keys
    assert all(("_w3j_" in k) or ("output_mask" in k) for k in res.missing_keys)
    own = model.state_dict()
    for k, v in g["state_dict"].items():
        assert own[k].shape == v.shape, k
    # layer structure of the reference: conv outputs 12x0e+4x1o+4x2e then 20x0e+4x1o+4x1e+4x2o+4x2e
    assert [l.equivariant_nonlin.plan.d_in for l in model.nequip.conv_layers] == [44, 84]
    assert model.nequip.out_irreps == ["4x0e+4x1o+4x2e", "4x0e+4x1o+4x1e+4x2o+4x2e"]


def test_gate_plan_matches_oracle_gate_layout():
    """GatePlan's offsets reproduce the reference Gate's sort-cut: evaluate the plan in numpy against the oracle module"""
    from e3nn import o3
    from oracle.ref_nequip import _Gate
    from physicsml_b200.plans import SILU_2MOM, TANH_2MOM, GatePlan

    for scal, gates, gated in [("8x0o+8x0e", "8x0e+8x0e+8x0e+8x0e", "8x1o+8x1e+8x2o+8x2e"), ("4x0e", "4x0e+4x0e", "4x1o+4x2e"),
                               ("3x0e", "", "")]:
        ref = _Gate(o3.Irreps(scal), o3.Irreps(gates), o3.Irreps(gated))
        plan = GatePlan(scal, gates, gated)
        assert plan.d_in == ref.irreps_in.dim and plan.d_out == ref.irreps_out.dim
        x = torch.randn(5, plan.d_in, generator=torch.Generator().manual_seed(1))
        y_ref = ref(x)
        st = plan.c_plan
        y = torch.zeros(5, plan.d_out)
        act = lambda kind, c, t: c * (torch.nn.functional.silu(t) if kind == 0 else torch.tanh(t))  # noqa: E731
        for i in range(st.n_scal):
            m = st.s_mul[i]
            y[:, st.s_out[i]:st.s_out[i] + m] = act(st.s_kind[i], st.s_c[i], x[:, st.s_in[i]:st.s_in[i] + m])
        for i in range(st.n_gated):
            m, d = st.g_mul[i], st.g_d[i]
            gate = act(st.g_kind[i], st.g_c[i], x[:, st.g_gate[i]:st.g_gate[i] + m])
            blk = x[:, st.g_x[i]:st.g_x[i] + m * d].reshape(5, m, d) * gate[:, :, None]
            y[:, st.g_out[i]:st.g_out[i] + m * d] = blk.reshape(5, m * d)
        assert (y - y_ref).abs().max() < 1e-6
        assert abs(SILU_2MOM - 1.6791767923989418) < 1e-12 and abs(TANH_2MOM - 1.5937334472592692) < 1e-12


def test_every_pdl_launched_kernel_waits_for_its_predecessors():
    """a kernel launched with the programmatic-dependent-launch attribute (pml::launch_pdl) may become resident while the
    previous kernel of the stream still runs: it MUST start with pdl_enter() (griddepcontrol.wait) or it reads stale data"""
    import glob
    import os
    import re

    csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "physicsml_b200", "csrc")
    launched, guarded = set(), set()
    for path in glob.glob(os.path.join(csrc, "*.cu")):
        src = open(path).read()
        for m in re.finditer(r"launch_pdl\(\s*(?:pml::)?(?:wide::)?([A-Za-z_0-9]+)", src):
            launched.add(m.group(1))
        for m in re.finditer(r"__global__", src):
            # declaration text up to the body's '{' at parenthesis depth 0; the kernel's name is the last identifier that
            # opens a parenthesis at depth 0 (attributes such as __launch_bounds__(...) come before it)
            i, depth, name, j = m.end(), 0, None, -1
            while i < len(src):
                c = src[i]
                if c == "(":
                    if depth == 0:
                        ident = re.search(r"([A-Za-z_0-9]+)\s*$", src[m.end():i])
                        name = ident.group(1) if ident else name
                    depth += 1
                elif c == ")":
                    depth -= 1
                elif c == ";" and depth == 0:
                    break
                elif c == "{" and depth == 0:
                    j = i
                    break
                i += 1
            if j >= 0 and name and src[j + 1 : j + 80].lstrip().startswith("pdl_enter();"):
                guarded.add(name)
    assert launched, "no launch_pdl call sites found"
    assert launched <= guarded, f"launched with the PDL attribute but without pdl_enter(): {sorted(launched - guarded)}"
