"""GPU parity of every custom op (called through the C ABI) against the oracle's dense PyTorch restatement of the
same op, evaluated on the same device in fp32 (and fp64 where cancellation matters).  Tolerances are stated inline."""
import math

import pytest
import torch

import helpers  # noqa: F401  (sets sys.path for the oracle shims)

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


def _edges(N, deg, gen, device):
    snd = torch.randint(0, N, (N * deg,), generator=gen)
    rcv = torch.randint(0, N, (N * deg,), generator=gen)
    return torch.stack([snd, rcv]).to(device)


# --------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("lmax", [1, 2, 3])
@pytest.mark.parametrize("pbc", [False, True])
def test_edge_featurise(cuda, lmax, pbc):
    from e3nn import o3
    from oracle.ref_model import _Radial, edge_vectors
    from physicsml_b200 import ops

    g = torch.Generator().manual_seed(lmax + 10 * pbc)
    N, E = 50, 700
    pos = (torch.rand(N, 3, generator=g) * 6).to(cuda)
    ei = _edges(N, 14, g, cuda)
    ei = ei[:, ei[0] != ei[1]]
    cell = shift = None
    if pbc:
        cell = (torch.eye(3) * 6 + 0.3 * torch.rand(3, 3, generator=g)).to(cuda)
        shift = torch.randint(-1, 2, (ei.shape[1], 3), generator=g).float().to(cuda)
    rc, p, nb = 5.0, 5, 8
    rad = _Radial(rc, nb, p).to(cuda)
    sh = o3.SphericalHarmonics(o3.Irreps.spherical_harmonics(lmax), normalize=True, normalization="component")
    plan = ops.EdgePlan(ei, N)

    pos_r = pos.clone().requires_grad_(True)
    d, r = edge_vectors(pos_r, ei, cell, shift)
    Yr, Br = sh(r), rad(d)
    pos_t = pos.clone().requires_grad_(True)
    Y, B = ops.edge_featurise(pos_t, plan.snd, plan.rcv, shift, cell, rad.bessel_fn.bessel_weights, rc, rad.bessel_fn.pref, 1.0, p, lmax)
    assert (Y - Yr).abs().max() < 2e-6  # fp32 polynomial evaluation, |Y| <= sqrt(7)
    assert (B - Br).abs().max() < 2e-6 * max(1.0, Br.abs().max().item())
    gY, gB = torch.randn(Y.shape, generator=g).to(cuda), torch.randn(B.shape, generator=g).to(cuda)
    (gr,) = torch.autograd.grad([Yr, Br], [pos_r], [gY, gB])
    gY_t, gB_t = gY.clone().requires_grad_(True), gB.clone().requires_grad_(True)
    (gt,) = torch.autograd.grad([Y, B], [pos_t], [gY_t, gB_t], create_graph=True)
    assert _rel(gt, gr) < 2e-5
    # tangent map = transpose of the cotangent map: <v, J^T g> == <J v, g>
    v = torch.randn(N, 3, generator=g).to(cuda)
    tY, tB = torch.autograd.grad([gt], [gY_t, gB_t], [v])
    lhs = (v * gt.detach()).sum().item()
    rhs = ((tY * gY).sum() + (tB * gB).sum()).item()
    assert abs(lhs - rhs) < 1e-3 * max(1.0, abs(lhs))


# --------------------------------------------------------------------------------------------------------------------
TP_CASES = [
    ("4x0e", 2, 4), ("4x0e+4x1o", 2, 4), ("32x0e+32x1o", 2, 32), ("8x0e+8x1o", 3, 8),
    ("128x0e+128x1o", 3, 128), ("40x0e", 3, 40), ("16x0e+16x1o+16x2e", 3, 16),
    ("6x0e+6x1o", 2, 6),  # NP*C = 42 floats: rows are not 16-byte multiples -> register-streaming kernels
]


def _ref_tp(node_irreps, lmax, C):
    from e3nn import o3
    from oracle.ref_model import uvu_instructions

    node = o3.Irreps(node_irreps)
    sh = o3.Irreps.spherical_harmonics(lmax)
    target = (sh * C).sort().irreps.simplify()
    tp_irreps, ins = uvu_instructions(node, sh, target)
    return o3.TensorProduct(node, sh, tp_irreps, instructions=ins, shared_weights=False, internal_weights=False), target


@pytest.mark.parametrize("node_irreps,lmax,C", TP_CASES)
def test_tp_scatter(cuda, node_irreps, lmax, C):
    from physicsml_b200 import ops
    from physicsml_b200.plans import TPPlan

    g = torch.Generator().manual_seed(7)
    N, deg = 37, 9
    ei = _edges(N, deg, g, cuda)
    E = ei.shape[1]
    tp_ref, target = _ref_tp(node_irreps, lmax, C)
    tp_ref = tp_ref.to(cuda)
    plan = TPPlan(node_irreps, lmax, str(target))
    h = torch.randn(N, tp_ref.irreps_in1.dim, generator=g).to(cuda).requires_grad_(True)
    Y = torch.randn(E, (lmax + 1) ** 2, generator=g).to(cuda).requires_grad_(True)
    R = torch.randn(E, tp_ref.weight_numel, generator=g).to(cuda).requires_grad_(True)
    assert tp_ref.weight_numel == plan.NP * C

    def ref(h, Y, R):
        m = tp_ref(h[ei[0]], Y, R)
        return torch.zeros(N, m.shape[1], device=cuda, dtype=m.dtype).index_add_(0, ei[1], m)

    ep = ops.EdgePlan(ei, N)
    out = ops.tp_scatter(h, Y, R, ep.gather, ep.rowptr, ep.perm, plan.handle)
    out_r = ref(h, Y, R)
    assert _rel(out, out_r) < 1e-5
    w = torch.randn(out.shape, generator=g).to(cuda).requires_grad_(True)
    gs = torch.autograd.grad([out], [h, Y, R], [w], create_graph=True)
    gr = torch.autograd.grad([out_r], [h, Y, R], [w], create_graph=True)
    for a, b in zip(gs, gr):
        assert _rel(a, b) < 1e-5
    # second order: differentiate <v, grads> w.r.t. everything
    vs = [torch.randn(a.shape, generator=g).to(cuda) for a in gs]
    s = sum((a * v).sum() for a, v in zip(gs, vs))
    sr = sum((a * v).sum() for a, v in zip(gr, vs))
    g2 = torch.autograd.grad(s, [h, Y, R, w])
    g2r = torch.autograd.grad(sr, [h, Y, R, w])
    for a, b in zip(g2, g2r):
        assert _rel(a, b) < 2e-5


def test_tp_second_order_fused_r_cotangent(cuda):
    """force-matching shape of the graph: a loss of the forward's output AND of its cotangents.  R then gets one cotangent
    from the second-order node and one from the forward's node; the second is accumulated onto the first inside the kernel
    (ops._PENDING_DR).  Same gradients as with the sum left to the engine (PML_FUSE_DR=0 semantics) and as e3nn; a pass
    restricted to inputs that need neither leaves nothing parked for the next one."""
    from physicsml_b200 import ops
    from physicsml_b200.plans import TPPlan

    g = torch.Generator().manual_seed(17)
    node_irreps, lmax, C, N, deg = "32x0e+32x1o", 2, 32, 41, 7
    ei = _edges(N, deg, g, cuda)
    E = ei.shape[1]
    tp_ref, target = _ref_tp(node_irreps, lmax, C)
    tp_ref = tp_ref.to(cuda)
    plan = TPPlan(node_irreps, lmax, str(target))
    h = torch.randn(N, tp_ref.irreps_in1.dim, generator=g).to(cuda).requires_grad_(True)
    Y = torch.randn(E, (lmax + 1) ** 2, generator=g).to(cuda).requires_grad_(True)
    R = torch.randn(E, tp_ref.weight_numel, generator=g).to(cuda).requires_grad_(True)
    ep = ops.EdgePlan(ei, N)
    vs = None

    def loss_of(fn):
        nonlocal vs
        out = fn(h, Y, R)
        w = torch.tanh(out)  # the cotangent depends on the output, as dE/d(out) does in the models
        gs = torch.autograd.grad([out], [h, Y, R], [w], create_graph=True)
        if vs is None:
            vs = [torch.randn(a.shape, generator=g).to(cuda) for a in gs]
        return (out ** 2).sum() + sum((a * v).sum() for a, v in zip(gs, vs)), w

    def mine(h, Y, R):
        return ops.tp_scatter(h, Y, R, ep.gather, ep.rowptr, ep.perm, plan.handle)

    def ref(h, Y, R):
        m = tp_ref(h[ei[0]], Y, R)
        return torch.zeros(N, m.shape[1], device=cuda, dtype=m.dtype).index_add_(0, ei[1], m)

    l_ref, _ = loss_of(ref)
    want = torch.autograd.grad(l_ref, [h, Y, R])
    results = {}
    for fuse in (True, False):
        ops.FUSE_DR[0] = fuse
        try:
            l, w = loss_of(mine)
            results[fuse] = torch.autograd.grad(l, [h, Y, R], retain_graph=True)
            assert not ops._PENDING_DR
            again = torch.autograd.grad(l, [h, Y, R])
            for a, b in zip(results[fuse], again):
                assert torch.equal(a, b) or _rel(a, b) < 1e-6  # atomics: not bit-identical
        finally:
            ops.FUSE_DR[0] = True
    for a, b, c in zip(results[True], results[False], want):
        assert _rel(a, c) < 2e-5 and _rel(b, c) < 2e-5 and _rel(a, b) < 1e-5
    # a restricted pass: only the seed of the first-order cotangent is asked for, so the forward's node is not executed
    out = mine(h, Y, R)
    w = torch.randn(out.shape, generator=g).to(cuda).requires_grad_(True)
    gs = torch.autograd.grad([out], [h, Y, R], [w], create_graph=True)
    s = sum((a * v).sum() for a, v in zip(gs, vs))
    (gw,) = torch.autograd.grad(s, [w], retain_graph=True)
    assert not ops._PENDING_DR  # dropped at the end of the pass, nothing leaks into the next one
    out_r = ref(h, Y, R)
    gr = torch.autograd.grad([out_r], [h, Y, R], [w], create_graph=True)
    (gw_r,) = torch.autograd.grad(sum((a * v).sum() for a, v in zip(gr, vs)), [w])
    assert _rel(gw, gw_r) < 2e-5
    full = torch.autograd.grad(s, [h, Y, R, w])
    full_r = torch.autograd.grad(sum((a * v).sum() for a, v in zip(torch.autograd.grad([ref(h, Y, R)], [h, Y, R], [w], create_graph=True), vs)),
                                 [h, Y, R, w])
    for a, b in zip(full, full_r):
        assert _rel(a, b) < 2e-5


@pytest.mark.parametrize("node_irreps,lmax,C,N,deg", [("128x0e+128x1o", 3, 128, 300, 33), ("40x0e", 3, 40, 1000, 5),
                                                      ("32x0e+32x1o", 2, 32, 513, 17), ("16x0e+16x1o+16x2e", 3, 16, 64, 100)])
def test_tp_pipelined_matches_streaming(cuda, node_irreps, lmax, C, N, deg):
    """The TMA-pipelined persistent kernels (packed fp32x2 arithmetic, factored differently) and the register-streaming
    kernels agree to fp32 rounding (1e-5 of the largest entry).  Includes isolated nodes and a node range long enough
    that every persistent CTA owns several nodes."""
    from physicsml_b200 import _lib, ops
    from physicsml_b200.plans import TPPlan

    g = torch.Generator().manual_seed(11)
    ei = _edges(N, deg, g, cuda)
    ei = ei[:, (ei[1] % 7) != 3]  # nodes 3, 10, 17, ... receive nothing
    E = ei.shape[1]
    _, target = _ref_tp(node_irreps, lmax, C)
    plan = TPPlan(node_irreps, lmax, str(target))
    h = torch.randn(N, plan.C * plan.DIN, generator=g).to(cuda)
    Y = torch.randn(E, (lmax + 1) ** 2, generator=g).to(cuda)
    R = torch.randn(E, plan.NP * C, generator=g).to(cuda)
    w = torch.randn(N, plan.C * plan.DOUT, generator=g).to(cuda)
    ep = ops.EdgePlan(ei, N)
    res = {}
    for simple in (0, 1):
        old = _lib.lib().pml_tp_set_simple(simple)
        try:
            out = ops.tp_scatter(h, Y, R, ep.gather, ep.rowptr, ep.perm, plan.handle)
            grads = ops.tp_scatter_bwd(h, Y, R, w, ep.gather, ep.rowptr, ep.perm, plan.handle, True, True, True)
            only_r = ops.tp_scatter_bwd(h, Y, R, w, ep.gather, ep.rowptr, ep.perm, plan.handle, False, False, True)[2]
            torch.cuda.synchronize()
        finally:
            _lib.lib().pml_tp_set_simple(old)
        res[simple] = (out, grads, only_r)
    assert _rel(res[0][0], res[1][0]) < 1e-5
    assert res[0][0][3].abs().max() == 0 and res[0][0][10].abs().max() == 0
    assert _rel(res[0][1][2], res[1][1][2]) < 1e-5  # dR
    assert _rel(res[0][2], res[1][2]) < 1e-5  # dR alone (R itself is not read)
    assert _rel(res[0][1][0], res[1][1][0]) < 1e-5  # dh: fp32 atomics
    assert _rel(res[0][1][1], res[1][1][1]) < 1e-5  # dY


@pytest.mark.parametrize("node_irreps,lmax,C,N,deg", [("128x0e+128x1o", 3, 128, 150, 21), ("32x0e+32x1o", 2, 32, 70, 9),
                                                      ("6x0e+6x1o", 2, 6, 40, 5)])
def test_tp_component_major_message_and_linear(cuda, node_irreps, lmax, C, N, deg):
    """The layout MACE uses between K4 and its per-irrep linear: tp_scatter with out_layout="component" is the
    e3nn-layout result with every path block transposed ([C, 2l+1] -> [2l+1, C]; pipelined and register-streaming
    kernels), and IrrepLinear(in_layout="component") on it equals the irreps-layout linear on the e3nn row - values,
    first derivatives and the second-order terms force-matching training takes."""
    from e3nn import o3
    from physicsml_b200 import ops
    from physicsml_b200.mace import IrrepLinear
    from physicsml_b200.plans import TPPlan

    g = torch.Generator().manual_seed(21)
    ei = _edges(N, deg, g, cuda)
    E = ei.shape[1]
    _, target = _ref_tp(node_irreps, lmax, C)
    pe = TPPlan(node_irreps, lmax, str(target))
    pc = TPPlan(node_irreps, lmax, str(target), out_layout="component")
    d_in = sum(m * (2 * l + 1) for m, l, _ in pe.out_irreps)
    h = torch.randn(N, o3.Irreps(node_irreps).dim, generator=g).to(cuda).requires_grad_(True)
    Y = torch.randn(E, (lmax + 1) ** 2, generator=g).to(cuda).requires_grad_(True)
    R = torch.randn(E, pe.NP * C, generator=g).to(cuda).requires_grad_(True)
    ep = ops.EdgePlan(ei, N)

    def to_component(row):  # e3nn row -> component-major row
        out, off = [], 0
        for m, l, _ in pe.out_irreps:
            d = 2 * l + 1
            out.append(row[:, off:off + m * d].reshape(-1, m, d).transpose(1, 2).reshape(-1, m * d))
            off += m * d
        return torch.cat(out, dim=1)

    torch.manual_seed(0)
    lin_e = IrrepLinear(pe.out_irreps, str(target), scale=0.3).to(cuda)
    lin_c = IrrepLinear(pc.out_irreps, str(target), scale=0.3, in_layout="component").to(cuda)
    lin_c.load_state_dict(lin_e.state_dict())
    gy = torch.randn(N, lin_e.plan.d_out, generator=g).to(cuda).requires_grad_(True)
    v = [torch.randn(t.shape, generator=g).to(cuda) for t in (h, Y, R)]
    res = []
    for plan, lin in ((pe, lin_e), (pc, lin_c)):
        a = ops.tp_scatter(h, Y, R, ep.gather, ep.rowptr, ep.perm, plan.handle)
        assert a.shape == (N, d_in)
        y = lin(a)
        gs = torch.autograd.grad([y], [h, Y, R, lin.weight], [gy], create_graph=True)
        g2 = torch.autograd.grad(sum((t * w).sum() for t, w in zip(gs[:3], v)), [h, R, lin.weight, gy])
        res.append((a.detach(), y.detach(), [t.detach() for t in gs], [t.detach() for t in g2]))
    assert _rel(res[1][0], to_component(res[0][0])) < 1e-6  # same arithmetic, different addresses
    assert _rel(res[1][1], res[0][1]) < 1e-5
    for a, b in zip(res[1][2] + res[1][3], res[0][2] + res[0][3]):
        assert _rel(a, b) < 2e-5


def test_kernel_timing_hook_sees_the_edge_kernels(cuda):
    """bench.py's roofline comes from ops.KERNEL_TIMING["pml_tp_fwd"]: the hook must keep seeing the fused edge kernels
    whichever entry point (pml_tp_fwd / pml_tp_fwd_ex) the op calls"""
    from physicsml_b200 import ops
    from physicsml_b200.plans import TPPlan

    g = torch.Generator().manual_seed(2)
    N, C, lmax = 12, 8, 2
    ei = _edges(N, 4, g, cuda)
    _, target = _ref_tp("8x0e+8x1o", lmax, C)
    ep = ops.EdgePlan(ei, N)
    try:
        ops.enable_kernel_timing(["pml_tp_fwd", "pml_tp_bwd"])
        for layout in ("irreps", "component"):
            plan = TPPlan("8x0e+8x1o", lmax, str(target), out_layout=layout)
            h = torch.randn(N, 32, generator=g).to(cuda).requires_grad_(True)
            Y = torch.randn(ei.shape[1], 9, generator=g).to(cuda)
            R = torch.randn(ei.shape[1], plan.NP * C, generator=g).to(cuda)
            ops.tp_scatter(h, Y, R, ep.gather, ep.rowptr, ep.perm, plan.handle).sum().backward()
        torch.cuda.synchronize()
        assert len(ops.KERNEL_TIMING["pml_tp_fwd"]) == 2 and len(ops.KERNEL_TIMING["pml_tp_bwd"]) == 2
        assert all(a.elapsed_time(b) > 0 for _, a, b in ops.KERNEL_TIMING["pml_tp_fwd"])
    finally:
        ops.enable_kernel_timing([])


@pytest.mark.parametrize("E", [0, 1, 5])
def test_tp_scatter_degenerate_graphs(cuda, E):
    """no edges at all / a single edge / fewer edges than persistent CTAs: every node row is still written (zeros for
    nodes that receive nothing) and the cotangents have the right shapes and values"""
    from physicsml_b200 import ops
    from physicsml_b200.plans import TPPlan

    g = torch.Generator().manual_seed(5)
    N, C, lmax = 9, 32, 2
    tp_ref, target = _ref_tp("32x0e+32x1o", lmax, C)
    tp_ref = tp_ref.to(cuda)
    plan = TPPlan("32x0e+32x1o", lmax, str(target))
    ei = torch.stack([torch.randint(0, N, (E,), generator=g), torch.randint(0, N, (E,), generator=g)]).to(cuda)
    h = torch.randn(N, 4 * C, generator=g).to(cuda).requires_grad_(True)
    Y = torch.randn(E, 9, generator=g).to(cuda).requires_grad_(True)
    R = torch.randn(E, plan.NP * C, generator=g).to(cuda).requires_grad_(True)
    ep = ops.EdgePlan(ei, N)
    out = ops.tp_scatter(h, Y, R, ep.gather, ep.rowptr, ep.perm, plan.handle)
    m = tp_ref(h[ei[0]], Y, R)
    ref = torch.zeros(N, out.shape[1], device=cuda).index_add_(0, ei[1], m)
    assert out.shape == ref.shape and torch.isfinite(out).all()
    assert (out - ref).abs().max() <= 1e-5 * max(1.0, ref.abs().max().item())
    w = torch.randn(out.shape, generator=g).to(cuda)
    gs = torch.autograd.grad([out], [h, Y, R], [w])
    gr = torch.autograd.grad([ref], [h, Y, R], [w], allow_unused=True)
    for a, b in zip(gs, gr):
        b = torch.zeros_like(a) if b is None else b
        assert a.shape == b.shape
        if a.numel():
            assert (a - b).abs().max() <= 1e-5 * max(1.0, b.abs().max().item())


def test_tp_scatter_grouped_edges(cuda):
    """perm=None fast path (edges already grouped by receiver)"""
    from physicsml_b200 import ops
    from physicsml_b200.plans import TPPlan

    g = torch.Generator().manual_seed(3)
    N, C, lmax = 20, 8, 2
    ei = _edges(N, 6, g, cuda)
    ei = ei[:, torch.argsort(ei[1], stable=True)]
    tp_ref, target = _ref_tp("8x0e+8x1o", lmax, C)
    tp_ref = tp_ref.to(cuda)
    plan = TPPlan("8x0e+8x1o", lmax, str(target))
    E = ei.shape[1]
    h = torch.randn(N, 32, generator=g).to(cuda)
    Y = torch.randn(E, 9, generator=g).to(cuda)
    R = torch.randn(E, plan.NP * C, generator=g).to(cuda)
    ep = ops.EdgePlan(ei, N, assume_grouped=True)
    out = ops.tp_scatter(h, Y, R, ep.gather, ep.rowptr, None, plan.handle)
    m = tp_ref(h[ei[0]], Y, R)
    ref = torch.zeros(N, m.shape[1], device=cuda).index_add_(0, ei[1], m)
    assert _rel(out, ref) < 1e-5


# --------------------------------------------------------------------------------------------------------------------
LIN_CASES = [
    ("4x0e", "4x0e", "irreps"), ("8x0e+8x1o", "8x0e+8x1o", "irreps"), ("64x0e+96x1o+64x2e", "32x0e+32x1o+32x2e", "irreps"),
    ("64x0e+96x1o+64x2e", "32x0e+32x1o+32x2e", "channel"), ("32x0e+32x1o", "1x0e", "irreps"), ("8x0e", "70x0e", "irreps"),
    ("256x0e+384x1o+384x2e+256x3o", "128x0e+128x1o+128x2e+128x3o", "channel"),
]


@pytest.mark.parametrize("irr_in,irr_out,layout", LIN_CASES)
@pytest.mark.parametrize("N", [1, 77, 300])
def test_irrep_linear(cuda, irr_in, irr_out, layout, N):
    from e3nn import o3
    from oracle.ref_model import to_channel_major
    from physicsml_b200 import ops
    from physicsml_b200.plans import LinearPlan

    g = torch.Generator().manual_seed(11)
    ref = o3.Linear(o3.Irreps(irr_in), o3.Irreps(irr_out)).to(cuda)
    plan = LinearPlan(irr_in, irr_out, scale=0.37, out_layout=layout)
    assert plan.weight_numel == ref.weight_numel
    x = torch.randn(N, ref.irreps_in.dim, generator=g).to(cuda).requires_grad_(True)
    w = ref.weight.detach().clone().requires_grad_(True)
    y = ops.irrep_linear(x, w, plan.handle)
    ref.weight.data.copy_(w.data)
    yr = ref(x) * 0.37
    if layout == "channel":
        yr = to_channel_major(yr, ref.irreps_out)
    assert _rel(y, yr) < 1e-5
    gy = torch.randn(y.shape, generator=g).to(cuda).requires_grad_(True)
    gs = torch.autograd.grad([y], [x, w], [gy], create_graph=True)
    gr = torch.autograd.grad([yr], [x, ref.weight], [gy], create_graph=True)
    for a, b in zip(gs, gr):
        assert _rel(a, b) < 2e-5
    vs = [torch.randn(a.shape, generator=g).to(cuda) for a in gs]
    g2 = torch.autograd.grad(sum((a * v).sum() for a, v in zip(gs, vs)), [x, w, gy])
    g2r = torch.autograd.grad(sum((a * v).sum() for a, v in zip(gr, vs)), [x, ref.weight, gy])
    for a, b in zip(g2, g2r):
        assert _rel(a, b) < 2e-5


@pytest.mark.parametrize("K,N,rows", [(64, 1280, 5000), (64, 512, 4097), (64, 1280, 20033), (32, 256, 4500), (8, 320, 4100)])
def test_wide_radial_layer_matches_fp64_and_general_kernel(cuda, K, N, rows):
    """K3w (csrc/gemm_wide.cu): the streaming kernels that take the wide radial-MLP output layer on edge-sized inputs
    (ragged last row block, K < 64 zero-padded, N not a multiple of 128) against fp64 and against the general grouped
    tensor-core kernel; first and second derivatives through the op's autograd formulas."""
    from physicsml_b200 import ops
    from physicsml_b200.plans import LinearPlan

    g = torch.Generator().manual_seed(3)
    plan = LinearPlan(f"{K}x0e", f"{N}x0e", scale=1.7)
    assert plan.wide == (K, N, plan.alpha[0])
    x = torch.randn(rows, K, generator=g).to(cuda).requires_grad_(True)
    w = torch.randn(K * N, generator=g).to(cuda).requires_grad_(True)
    gy = torch.randn(rows, N, generator=g).to(cuda).requires_grad_(True)
    vx = torch.randn(rows, K, generator=g).to(cuda)

    def run():
        y = ops.irrep_linear(x, w, plan.handle)
        dx, dw = torch.autograd.grad([y], [x, w], [gy], create_graph=True)
        d2 = torch.autograd.grad((dx * vx).sum(), [w, gy])
        return [y, dx, dw, *d2]

    old = ops.WIDE_MIN_ROWS[0]
    try:
        ops.WIDE_MIN_ROWS[0] = 4096
        assert ops._wide_ok(plan, rows)
        wide = [t.detach().clone() for t in run()]
        ops.WIDE_MIN_ROWS[0] = 1 << 30
        general = [t.detach().clone() for t in run()]
    finally:
        ops.WIDE_MIN_ROWS[0] = old
    a = 1.7 / K ** 0.5
    W = w.detach().double().view(K, N)
    ref = [a * x.detach().double() @ W, a * gy.detach().double() @ W.t(), a * (x.detach().double().t() @ gy.detach().double()).reshape(-1),
           a * (vx.double().t() @ gy.detach().double()).reshape(-1), a * vx.double() @ W]
    for got_w, got_g, want in zip(wide, general, ref):
        assert _rel(got_w.double(), want) < 5e-6
        assert _rel(got_g.double(), want) < 5e-6


@pytest.mark.parametrize("K,N,rows", [(64, 64, 5000), (32, 64, 4133), (64, 32, 9001)])
def test_narrow_radial_layers_through_the_streaming_kernel(cuda, K, N, rows):
    """the hidden layers of the radial MLP on very large edge sets: y = x W and dx = g W^T both run the streaming
    (TMEM-operand) dx kernel with the weight image packed in the matching orientation"""
    from physicsml_b200 import ops
    from physicsml_b200.plans import LinearPlan

    g = torch.Generator().manual_seed(5)
    plan = LinearPlan(f"{K}x0e", f"{N}x0e", scale=0.9)
    assert plan.narrow == (K, N, plan.alpha[0])
    x = torch.randn(rows, K, generator=g).to(cuda)
    w = torch.randn(K * N, generator=g).to(cuda)
    gy = torch.randn(rows, N, generator=g).to(cuda)
    old = ops.NARROW_MIN_ROWS[0]
    try:
        ops.NARROW_MIN_ROWS[0] = 1024
        assert ops._narrow_ok(plan, rows)
        y, dx = ops.irrep_linear(x, w, plan.handle), ops.irrep_linear_bwd_x(gy, w, plan.handle)
        ops.NARROW_MIN_ROWS[0] = 1 << 30
        y0, dx0 = ops.irrep_linear(x, w, plan.handle), ops.irrep_linear_bwd_x(gy, w, plan.handle)
    finally:
        ops.NARROW_MIN_ROWS[0] = old
    a = 0.9 / K ** 0.5
    W = w.double().view(K, N)
    for got, got0, want in ((y, y0, a * x.double() @ W), (dx, dx0, a * gy.double() @ W.t())):
        assert _rel(got.double(), want) < 5e-6 and _rel(got0.double(), want) < 5e-6


@pytest.mark.parametrize("irr_in,irr_out,layout", [
    ("8x0e+8x1o+8x2e", "8x0e+8x1o+8x2e", "irreps"), ("8x0e+8x1o+8x2e", "8x0e+8x1o+8x2e", "channel"),
    ("32x0e+32x1o", "32x0e", "irreps"), ("128x0e+128x1o+128x2e+128x3o", "128x0e+128x1o+128x2e+128x3o", "channel"),
])
@pytest.mark.parametrize("onehot", [True, False])
def test_element_linear(cuda, irr_in, irr_out, layout, onehot):
    from e3nn import o3
    from oracle.ref_model import to_channel_major
    from physicsml_b200 import ops
    from physicsml_b200.plans import ElementLinearPlan

    g = torch.Generator().manual_seed(5)
    Z, N = 5, 61
    ref = o3.FullyConnectedTensorProduct(o3.Irreps(irr_in), o3.Irreps(f"{Z}x0e"), o3.Irreps(irr_out)).to(cuda)
    plan = ElementLinearPlan(irr_in, Z, irr_out, out_layout=layout)
    assert plan.weight_numel == ref.weight_numel
    x = torch.randn(N, ref.irreps_in1.dim, generator=g).to(cuda).requires_grad_(True)
    if onehot:
        attrs = torch.nn.functional.one_hot(torch.randint(0, Z, (N,), generator=g), Z).float().to(cuda)
    else:
        attrs = torch.randn(N, Z, generator=g).to(cuda)
    w = ref.weight.detach().clone().requires_grad_(True)
    y = ops.element_linear(x, attrs, w, plan.handle)
    yr = ref(x, attrs)
    if layout == "channel":
        yr = to_channel_major(yr, ref.irreps_out)
    assert _rel(y, yr) < 1e-5
    gy = torch.randn(y.shape, generator=g).to(cuda).requires_grad_(True)
    gs = torch.autograd.grad([y], [x, w], [gy], create_graph=True)
    gr = torch.autograd.grad([yr], [x, ref.weight], [gy], create_graph=True)
    for a, b in zip(gs, gr):
        assert _rel(a, b) < 2e-5
    vs = [torch.randn(a.shape, generator=g).to(cuda) for a in gs]
    g2 = torch.autograd.grad(sum((a * v).sum() for a, v in zip(gs, vs)), [x, w, gy])
    g2r = torch.autograd.grad(sum((a * v).sum() for a, v in zip(gr, vs)), [x, ref.weight, gy])
    for a, b in zip(g2, g2r):
        assert _rel(a, b) < 2e-5


def test_species_plan_orders_nodes_by_element(cuda):
    """pml_species_order: node ids grouped by element, ascending inside a group; group starts; the not-one-hot flag"""
    from physicsml_b200 import ops

    g = torch.Generator().manual_seed(3)
    for N, Z in ((1, 3), (700, 7), (5000, 4), (2049, 128)):
        z = torch.randint(0, Z, (N,), generator=g)
        if Z == 7:
            z[z == 2] = 3  # an element that does not occur
        attrs = torch.nn.functional.one_hot(z, Z).float().to(cuda)
        sp = ops.SpeciesPlan(attrs)
        assert sp.one_hot()
        want = torch.sort(z, stable=True).indices
        assert torch.equal(sp.order.cpu().long(), want)
        counts = torch.bincount(z, minlength=Z)
        assert sp.seg[: Z + 1].cpu().tolist() == [0] + torch.cumsum(counts, 0).tolist()
    attrs[5, :] = 0.5
    sp = ops.SpeciesPlan(attrs)
    assert not sp.one_hot()
    assert sorted(sp.order.cpu().tolist()) == list(range(attrs.shape[0]))  # still a permutation (rejected rows at the tail)
    assert sp.order[-1].item() == 5


@pytest.mark.parametrize("irr_in,irr_out,layout,Z", [
    ("128x0e+128x1o+128x2e+128x3o", "128x0e+128x1o+128x2e+128x3o", "channel", 10),
    ("64x0e+64x1o+64x2e", "64x0e+64x1e+64x1o+64x2e", "irreps", 5),   # an output block nobody feeds
    ("64x0e+64x1o+64x1e", "64x0e+64x1o", "irreps", 3),               # an input block that feeds nothing
    ("256x0e+256x1o+256x2e", "256x0e+256x1o+256x2e", "channel", 4),  # reductions longer than 128: promoting kernel
])
def test_element_linear_grouped_by_element(cuda, irr_in, irr_out, layout, Z):
    """the tensor-core route of element_linear (nodes grouped by element, ops.SpeciesPlan) against e3nn's
    FullyConnectedTensorProduct and against the FFMA kernels: values, first and second order"""
    from e3nn import o3
    from oracle.ref_model import to_channel_major
    from physicsml_b200 import ops
    from physicsml_b200.plans import ElementLinearPlan

    g = torch.Generator().manual_seed(7)
    N = 777
    ref = o3.FullyConnectedTensorProduct(o3.Irreps(irr_in), o3.Irreps(f"{Z}x0e"), o3.Irreps(irr_out)).to(cuda)
    plan = ElementLinearPlan(irr_in, Z, irr_out, out_layout=layout)
    assert plan.seg_ok
    z = torch.randint(0, Z, (N,), generator=g)
    z[z == 1] = 0  # element 1 does not occur in the batch
    attrs = torch.nn.functional.one_hot(z, Z).float().to(cuda)
    sp = ops.species_plan_for({"node_attrs": attrs}, [plan])
    assert sp is not None
    x = torch.randn(N, ref.irreps_in1.dim, generator=g).to(cuda).requires_grad_(True)
    w = ref.weight.detach().clone().requires_grad_(True)
    gemms, real_gemm = [], ops._gemm
    ops._gemm = lambda name, *a: (gemms.append(name), real_gemm(name, *a))[1]
    try:
        y = ops.element_linear(x, attrs, w, plan.handle, sp.handle)
    finally:
        ops._gemm = real_gemm
    assert gemms == ["pml_gemm_grouped_tc"]  # the grouped tensor-core route ran, not the FFMA kernels
    y0 = ops.element_linear(x, attrs, w, plan.handle)
    yr = ref(x, attrs)
    if layout == "channel":
        yr = to_channel_major(yr, ref.irreps_out)
    assert _rel(y, yr) < 1e-5 and _rel(y0, yr) < 1e-5
    gy = torch.randn(y.shape, generator=g).to(cuda).requires_grad_(True)
    gs = torch.autograd.grad([y], [x, w], [gy], create_graph=True)
    gr = torch.autograd.grad([yr], [x, ref.weight], [gy], create_graph=True)
    for a, b in zip(gs, gr):
        assert _rel(a, b) < 2e-5
    vs = [torch.randn(a.shape, generator=g).to(cuda) for a in gs]
    g2 = torch.autograd.grad(sum((a * v).sum() for a, v in zip(gs, vs)), [x, w, gy])
    g2r = torch.autograd.grad(sum((a * v).sum() for a, v in zip(gr, vs)), [x, ref.weight, gy])
    for a, b in zip(g2, g2r):
        assert _rel(a, b) < 2e-5
    # a batch that is not one-hot must not silently use the wrong weights: no plan from the host check ...
    bad = attrs.clone()
    bad[3] = 0.25
    assert ops.species_plan_for({"node_attrs": bad}, [plan]) is None
    # ... and NaN from a plan built without it (what a replayed CUDA graph would do)
    spb = ops.SpeciesPlan(bad)
    assert torch.isnan(ops.element_linear(x, bad, w, plan.handle, spb.handle)).any()


# --------------------------------------------------------------------------------------------------------------------
SYM_CASES = [(2, "4x0e+4x1o", 2, 4), (2, "4x0e", 2, 4), (2, "32x0e+32x1o", 3, 32), (3, "16x0e+16x1o", 3, 16),
             (3, "128x0e", 3, 128), (3, "8x0e+8x1o+8x2e", 3, 8)]


@pytest.mark.parametrize("lmax,hidden,corr,C", SYM_CASES)
@pytest.mark.parametrize("onehot", [True, False])
def test_sym_contract(cuda, lmax, hidden, corr, C, onehot):
    from e3nn import o3
    from oracle.ref_model import _SymContraction
    from physicsml_b200.mace import SymmetricContraction

    g = torch.Generator().manual_seed(13)
    Z, N = 4, 23
    inter = (o3.Irreps.spherical_harmonics(lmax) * C).sort().irreps.simplify()
    torch.manual_seed(0)
    ref = _SymContraction(inter, o3.Irreps(hidden), corr, Z).to(cuda)
    mine = SymmetricContraction(str(inter), hidden, corr, Z).to(cuda)
    mine.load_state_dict(ref.state_dict())
    S = (lmax + 1) ** 2
    x = torch.randn(N, C, S, generator=g).to(cuda).requires_grad_(True)
    if onehot:
        attrs = torch.nn.functional.one_hot(torch.randint(0, Z, (N,), generator=g), Z).float().to(cuda)
    else:
        attrs = torch.randn(N, Z, generator=g).to(cuda)
    y, yr = mine(x, attrs), ref(x, attrs)
    assert _rel(y, yr) < 2e-5
    pm = [p for p in mine.parameters() if p.requires_grad]
    pr = [p for p in ref.parameters() if p.requires_grad]
    gy = torch.randn(y.shape, generator=g).to(cuda).requires_grad_(True)
    gs = torch.autograd.grad([y], [x] + pm, [gy], create_graph=True)
    gr = torch.autograd.grad([yr], [x] + pr, [gy], create_graph=True)
    for a, b in zip(gs, gr):
        assert _rel(a, b) < 5e-5
    # second order through the x-cotangent (what force-matching training differentiates)
    v = torch.randn(x.shape, generator=g).to(cuda)
    g2 = torch.autograd.grad((gs[0] * v).sum(), [x, gy] + pm)
    g2r = torch.autograd.grad((gr[0] * v).sum(), [x, gy] + pr)
    for a, b in zip(g2, g2r):
        assert _rel(a, b) < 1e-4


def test_act_and_pool(cuda):
    from physicsml_b200 import ops

    g = torch.Generator().manual_seed(1)
    x = (3 * torch.randn(1000, generator=g)).to(cuda).requires_grad_(True)
    c = 1.6791767923989418
    y = ops.act(x, 0, c)
    yr = c * torch.nn.functional.silu(x)
    assert (y - yr).abs().max() < 1e-5
    gy = torch.randn(1000, generator=g).to(cuda).requires_grad_(True)
    (a,) = torch.autograd.grad(y, x, gy, create_graph=True)
    (b,) = torch.autograd.grad(yr, x, gy, create_graph=True)
    assert (a - b).abs().max() < 1e-5
    v = torch.randn(1000, generator=g).to(cuda)
    a2 = torch.autograd.grad((a * v).sum(), [x, gy])
    b2 = torch.autograd.grad((b * v).sum(), [x, gy])
    for p, q in zip(a2, b2):
        assert (p - q).abs().max() < 1e-5
    batch = torch.sort(torch.randint(0, 7, (200,), generator=g)).values.to(cuda)
    f = torch.randn(200, 3, generator=g).to(cuda).requires_grad_(True)
    out = ops.pool_sum(f, batch, 7)
    ref = torch.zeros(7, 3, device=cuda).index_add_(0, batch, f)
    assert (out - ref).abs().max() < 1e-5
    (gf,) = torch.autograd.grad(out.pow(2).sum(), f)
    (gfr,) = torch.autograd.grad(ref.pow(2).sum(), f)
    assert (gf - gfr).abs().max() < 1e-4


def test_cpu_tensors_fail_loudly():
    from physicsml_b200 import ops

    with pytest.raises((NotImplementedError, RuntimeError)):
        ops.act(torch.randn(4), 0, 1.0)


@pytest.mark.parametrize("irr_in,irr_out", [("8x0e", "64x0e"), ("64x0e", "64x0e"), ("64x0e", "1280x0e"), ("64x0e", "96x0e"),
                                            ("256x0e+384x1o+384x2e+256x3o", "128x0e+128x1o+128x2e+128x3o"),
                                            ("40x0e+40x1o", "24x0e+24x1o")])
@pytest.mark.parametrize("N", [64, 130, 4100])
def test_tensor_core_gemm_matches_ffma(cuda, irr_in, irr_out, N):
    """tcgen05 3xTF32 kernel and the FFMA kernel on the same problem tables (fwd, dx, dW), both against an fp64
    evaluation of the oracle's Linear: the tensor-core path must stay at fp32-level accuracy (<= 2e-5 of the largest
    output for reductions up to K = 1280; the FFMA kernel itself sits at ~2e-6 there)."""
    from e3nn import o3
    from physicsml_b200 import ops
    from physicsml_b200.plans import LinearPlan

    g = torch.Generator().manual_seed(17)
    plan = LinearPlan(irr_in, irr_out, scale=0.7)
    x = torch.randn(N, plan.d_in, generator=g).to(cuda)
    w = torch.randn(plan.weight_numel, generator=g).to(cuda)
    gy = torch.randn(N, plan.d_out, generator=g).to(cuda)
    ref = o3.Linear(o3.Irreps(irr_in), o3.Irreps(irr_out)).to(cuda).double()
    ref.weight.data.copy_(w.double())
    x64 = x.double().requires_grad_(True)
    y64 = ref(x64) * 0.7
    dx64, dw64 = torch.autograd.grad([y64], [x64, ref.weight], [gy.double()])
    exact = (y64.detach(), dx64, dw64)
    res = {}
    old_rows = ops.TC_MIN_ROWS[0]
    ops.TC_MIN_ROWS[0] = 1
    try:
        for mode in ("ffma", "tc"):
            ops.GEMM_MODE[0] = mode
            res[mode] = (ops.irrep_linear(x, w, plan.handle), ops.irrep_linear_bwd_x(gy, w, plan.handle),
                         ops.irrep_linear_bwd_w(x, gy, plan.handle))
    finally:
        ops.GEMM_MODE[0] = "tc"
        ops.TC_MIN_ROWS[0] = old_rows
    for a, b, e, name in zip(res["tc"], res["ffma"], exact, ("fwd", "dx", "dw")):
        err_tc, err_ff = _rel(a.double(), e), _rel(b.double(), e)
        print(name, irr_in, irr_out, N, "rel err vs fp64: tc", err_tc, "ffma", err_ff)
        assert err_tc < 2e-5, (name, err_tc)
        assert err_ff < 2e-5, (name, err_ff)


def test_fused_adam_matches_torch(cuda):
    """FusedAdam (one kernel for all tensors, device-side step counter) against torch.optim.Adam(amsgrad=True,
    weight_decay) over several steps; also checks that a changed gradient pointer set rebuilds the tables"""
    from physicsml_b200.optim import FusedAdam

    g = torch.Generator().manual_seed(3)
    shapes = [(5,), (64, 12), (4097,), (10, 23, 128), (1,)]
    p_ref = [torch.randn(s, generator=g).to(cuda).requires_grad_(True) for s in shapes]
    p_mine = [p.detach().clone().requires_grad_(True) for p in p_ref]
    kw = dict(lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=5e-7, amsgrad=True)
    o_ref, o_mine = torch.optim.Adam(p_ref, **kw), FusedAdam(p_mine, **kw)
    for step in range(6):
        grads = [torch.randn(s, generator=g).to(cuda) for s in shapes]
        for a, b, gr in zip(p_ref, p_mine, grads):
            a.grad = gr.clone()
            b.grad = gr.clone()  # new tensors every step: the pointer table must follow
        o_ref.step()
        o_mine.step()
        for a, b in zip(p_ref, p_mine):
            assert _rel(b.detach(), a.detach()) < 1e-6, step
    assert float(o_mine.state[p_mine[0]]["step"]) == 6.0
    for a, b in zip(p_ref, p_mine):
        assert _rel(o_mine.state[b]["max_exp_avg_sq"], o_ref.state[a]["max_exp_avg_sq"]) < 1e-6


def test_fused_energy_force_mse(cuda):
    """the fused loss == MSELoss(E) + MSELoss(F) with F = -dE/dx, value and gradients"""
    from physicsml_b200 import ops

    g = torch.Generator().manual_seed(1)
    e, e_t = torch.randn(32, 1, generator=g), torch.randn(32, 1, generator=g)
    gp, f_t = torch.randn(1304, 3, generator=g), torch.randn(1304, 3, generator=g)
    a = [t.clone().to(cuda).requires_grad_(r) for t, r in ((e, True), (e_t, False), (gp, True), (f_t, False))]
    loss = ops.energy_force_mse(a[0], a[1], a[2], a[3], 1.0, 10.0)
    (loss * 3.0).backward()
    b = [t.clone().double().requires_grad_(r) for t, r in ((e, True), (e_t, False), (gp, True), (f_t, False))]
    ref = torch.nn.functional.mse_loss(b[0], b[1]) + 10.0 * torch.nn.functional.mse_loss(-b[2], b[3])
    (ref * 3.0).backward()
    assert abs(loss.item() - ref.item()) <= 2e-6 * abs(ref.item())
    assert (a[0].grad.cpu().double() - b[0].grad).abs().max().item() <= 1e-6 * b[0].grad.abs().max().item()
    assert (a[2].grad.cpu().double() - b[2].grad).abs().max().item() <= 1e-6 * b[2].grad.abs().max().item()
    again = ops.energy_force_mse(a[0], a[1], a[2], a[3], 1.0, 10.0)
    assert again.item() == loss.item()  # one CTA, fixed order: deterministic
