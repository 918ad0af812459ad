"""Host-side mirror of the reference's MACE module interface, running on the B200 kernels.

Class names, constructor arguments, ``forward`` signatures, the batch-dict contract and every state-dict key follow
``physicsml.models.mace.modules`` (reference ``src/physicsml/models/mace/modules/mace.py:18-260``, ``blocks.py:11-256``,
``symmetric_contraction.py:24-243``) so a reference checkpoint loads with ``strict=True`` and a caller written against
the reference (``PooledMACEModule.forward`` -> ``compute_forces_by_gradient``, ``lightning/module.py:28-48``) works
unchanged.  All arithmetic runs in the custom CUDA ops of :mod:`physicsml_b200.ops`; there is no eager fallback.
"""
from __future__ import annotations

import math
from typing import Any, Optional

import torch

from . import ops, so3
from .plans import ElementLinearPlan, LinearPlan, SymPlan, TPPlan

SILU_2MOM = 1.6791767923989418  # e3nn normalize2mom(SiLU) constant

# The radial MLPs run on a second stream next to the node-level chain (see MACE.forward) on graphs where the step is a
# chain of small kernels: config 2 5.53 -> 5.29 ms, config 1 0.61 -> 0.52 ms, config 5 22.9 -> 22.5 ms.  Off for large
# single structures, where every kernel fills the GPU and two streams only fight over L2 (config 4: 34.4 -> 44.4 ms).
# PML_SIDE_STREAM=0 disables it.
SIDE_STREAM = [__import__("os").environ.get("PML_SIDE_STREAM", "1") == "1"]
SIDE_STREAM_MIN_EDGES = 2048
SIDE_STREAM_MAX_EDGES = 262144
_SIDE: dict = {}


def side_stream_ok(edge_feats: torch.Tensor) -> bool:
    return bool(SIDE_STREAM[0]) and edge_feats.is_cuda and SIDE_STREAM_MIN_EDGES <= edge_feats.shape[0] <= SIDE_STREAM_MAX_EDGES


def on_side_stream(fn, *tensors):
    """run ``fn(*tensors)`` on the side stream -> (result, event); the consumer waits for the event on its own stream
    right before it reads the result (autograd runs the cotangent of `fn` on the side stream as well)"""
    main = torch.cuda.current_stream()
    side = _side_stream(tensors[0].device)
    side.wait_stream(main)
    with torch.cuda.stream(side):
        out = fn(*tensors)
        out.record_stream(main)
        ev = side.record_event()
    for t in tensors:
        t.record_stream(side)
    return out, ev


def radial_on_side_stream(mlps, edge_feats: torch.Tensor) -> list:
    """run the radial MLPs `mlps` (one per layer, all functions of the edge features only) on the side stream;
    -> [(R_i, event_i)]: the consumer waits for event_i on its own stream right before it reads R_i"""
    main = torch.cuda.current_stream()
    side = _side_stream(edge_feats.device)
    side.wait_stream(main)
    out = []
    with torch.cuda.stream(side):
        for mlp in mlps:
            r = mlp(edge_feats)
            r.record_stream(main)
            out.append((r, side.record_event()))
    edge_feats.record_stream(side)
    return out


def _side_stream(device) -> "torch.cuda.Stream":
    key = (device.index if device.index is not None else torch.cuda.current_device())
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=device)
    return _SIDE[key]


def _irreps(x) -> list:
    """accepts 'AxBe+...' strings, lists of (mul, l, p) and anything whose str() is an irreps string (e3nn Irreps)"""
    if isinstance(x, (list, tuple)) and (len(x) == 0 or isinstance(x[0], (list, tuple))):
        return so3.parse_irreps(x)
    return so3.parse_irreps(str(x))


class IrrepLinear(torch.nn.Module):
    """o3.Linear with internal shared weights (flat ``weight``; ``bias``/``output_mask`` kept for state-dict parity)."""

    def __init__(self, irreps_in, irreps_out, scale: float = 1.0, out_layout: str = "irreps",
                 in_layout: str = "irreps") -> None:
        super().__init__()
        self.plan = LinearPlan(irreps_in, irreps_out, scale=scale, out_layout=out_layout, in_layout=in_layout)
        self.weight = torch.nn.Parameter(torch.randn(self.plan.weight_numel))
        self.bias = torch.nn.Parameter(torch.zeros(0))
        mask = torch.zeros(self.plan.d_out)
        offs = so3.irreps_offsets(self.plan.irreps_out)
        for _, o, _ in self.plan.paths:
            m, l, _p = self.plan.irreps_out[o]
            mask[offs[o] : offs[o] + m * (2 * l + 1)] = 1.0
        self.register_buffer("output_mask", mask)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return ops.irrep_linear(x, self.weight, self.plan.handle)


class ElementLinear(torch.nn.Module):
    """o3.FullyConnectedTensorProduct(x, Zx0e) with internal weights."""

    def __init__(self, irreps_in, num_elements: int, irreps_out, out_layout: str = "irreps") -> None:
        super().__init__()
        self.plan = ElementLinearPlan(irreps_in, num_elements, irreps_out, out_layout=out_layout)
        self.weight = torch.nn.Parameter(torch.randn(self.plan.weight_numel))
        mask = torch.zeros(self.plan.d_out)
        self.register_buffer("output_mask", mask + 1.0 if not self.plan.has_unfed_output else mask)

    def forward(self, x: torch.Tensor, node_attrs: torch.Tensor, species: Optional["ops.SpeciesPlan"] = None) -> torch.Tensor:
        """species: the batch's nodes grouped by element (ops.species_plan_for): the contraction then runs as grouped
        tensor-core GEMMs over contiguous row ranges instead of the per-node FFMA kernels"""
        return ops.element_linear(x, node_attrs, self.weight, self.plan.handle, -1 if species is None else species.handle)


class _DenseLayer(torch.nn.Module):
    def __init__(self, h_in: int, h_out: int) -> None:
        super().__init__()
        self.weight = torch.nn.Parameter(torch.randn(h_in, h_out))
        self.plan = LinearPlan(f"{h_in}x0e", f"{h_out}x0e")  # alpha = 1/sqrt(h_in)


class RadialMLP(torch.nn.Module):
    """e3nn nn.FullyConnectedNet(hs, SiLU): x <- c*silu(x @ W / sqrt(h_in)) on hidden layers, linear last layer."""

    def __init__(self, hs) -> None:
        super().__init__()
        self.hs = list(hs)
        for i, (a, b) in enumerate(zip(self.hs, self.hs[1:])):
            setattr(self, f"layer{i}", _DenseLayer(a, b))
        self.n_layers = len(self.hs) - 1

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        for i in range(self.n_layers):
            layer = getattr(self, f"layer{i}")
            x = ops.irrep_linear(x, layer.weight.reshape(-1), layer.plan.handle)
            if i < self.n_layers - 1:
                x = ops.act(x, 0, SILU_2MOM)
        return x


class _TPBuffers(torch.nn.Module):
    """State-dict parity with o3.TensorProduct(shared_weights=False): empty ``weight``, ``output_mask`` and the 3j
    buffers e3nn's generated code stores for paths with l1,l2,l3 > 0.  Not used by the kernels (tables are baked in)."""

    def __init__(self, tp: TPPlan) -> None:
        super().__init__()
        self.register_buffer("weight", torch.zeros(0))
        self.register_buffer("output_mask", torch.ones(tp.C * tp.DOUT))
        self._compiled_main_left_right = torch.nn.Module()
        for q in tp.spec.paths:
            name = f"_w3j_{q.l1}_{q.l2}_{q.l3}"
            if min(q.l1, q.l2, q.l3) > 0 and not hasattr(self._compiled_main_left_right, name):
                self._compiled_main_left_right.register_buffer(
                    name, torch.tensor(so3.wigner_3j(q.l1, q.l2, q.l3), dtype=torch.get_default_dtype())
                )


class RadialEmbeddingBlock(torch.nn.Module):
    """Only carries ``bessel_fn.bessel_weights`` (reference blocks.py:31-51, radial.py:6-30); evaluated inside K2."""

    def __init__(self, r_max: float, num_bessel: int, num_polynomial_cutoff: int) -> None:
        super().__init__()
        self.bessel_fn = torch.nn.Module()
        self.bessel_fn.register_buffer("bessel_weights", math.pi * torch.arange(1, num_bessel + 1) / r_max)
        self.r_max, self.p, self.out_dim = float(r_max), int(num_polynomial_cutoff), num_bessel
        self.prefactor = math.sqrt(2.0 / r_max)


class InteractionBlock(torch.nn.Module):
    """reference blocks.py:124-237"""

    def __init__(self, node_feats_irreps, node_attrs_irreps, edge_attrs_irreps, edge_feats_irreps, interaction_irreps,
                 avg_num_neighbours: float, mix_with_node_attrs: bool = False) -> None:
        super().__init__()
        node = _irreps(node_feats_irreps)
        attrs = _irreps(node_attrs_irreps)
        sh = _irreps(edge_attrs_irreps)
        inter = _irreps(interaction_irreps)
        self.avg_num_neighbours = avg_num_neighbours
        lmax = max(l for _, l, _ in sh)
        self.lmax = lmax
        self.linear_node_feats = IrrepLinear(node, node)
        # the aggregated message only ever feeds `linear`: it is kept component-major (m*C + u inside every path
        # block), which gives the tensor-product kernels packed stores and the GEMM a unit-stride reduction index
        self.tp = TPPlan(node, lmax, inter, sh_parity=-1 if any(p == -1 for _, _, p in sh) else 1, out_layout="component")
        n_edge_feats = so3.irreps_dim(_irreps(edge_feats_irreps))
        self.net_R_channels = RadialMLP([n_edge_feats, 64, 64, 64, self.tp.NP * self.tp.C])
        self.conv_tp = _TPBuffers(self.tp)
        channel = all(m == inter[0][0] for m, _, _ in inter)
        tp_simplified = so3.simplify_irreps(self.tp.out_irreps)
        self.channels = inter[0][0]
        # '/ avg_num_neighbours' (blocks.py:229) is folded into the GEMM's alpha; the final reshape (blocks.py:235)
        # into its output strides (of `linear` for layers >= 1, of `mix_layer` for layer 0)
        # one input block per path (same flat weight layout as the simplified irreps: blocks of one irrep are adjacent)
        assert so3.simplify_irreps(self.tp.out_irreps) == tp_simplified
        self.linear = IrrepLinear(self.tp.out_irreps, inter, scale=1.0 / avg_num_neighbours,
                                  out_layout="irreps" if (mix_with_node_attrs or not channel) else "channel",
                                  in_layout="component")
        self.mix_layer = (
            ElementLinear(inter, so3.irreps_dim(attrs), inter, out_layout="channel" if channel else "irreps")
            if mix_with_node_attrs else None
        )
        self._channel = channel

    def forward(self, node_feats, node_attrs, edge_attrs, edge_feats, edge_index, edge_plan: Optional[ops.EdgePlan] = None,
                radial: Optional[torch.Tensor] = None, species: Optional["ops.SpeciesPlan"] = None):
        """radial: the block's radial-MLP output if the caller already computed it (MACE.forward runs the radial MLPs of all
        layers on a side stream: they depend on the edge features only)"""
        if edge_plan is None:
            edge_plan = ops.EdgePlan(edge_index, node_feats.shape[0], scatter_row=1)
        wh = self.linear_node_feats(node_feats)
        if radial is None:
            R = self.net_R_channels(edge_feats)
        else:
            R, ready = radial
            torch.cuda.current_stream().wait_event(ready)  # computed on the side stream: join right before its first consumer
        A = ops.tp_scatter(wh, edge_attrs, R, edge_plan.gather, edge_plan.rowptr, edge_plan.perm, self.tp.handle)
        A = self.linear(A)
        if self.mix_layer is not None:
            A = self.mix_layer(A, node_attrs, species)
        if not self._channel:
            A = ops.irreps_to_channel(A, self.channels, self.lmax, False)
        return A  # [N, C, S]


class Contraction(torch.nn.Module):
    """Parameter container for one output irrep (reference symmetric_contraction.py:81-243)."""

    def __init__(self, coupling, ir_out, corr: int, n_feat: int, n_elem: int) -> None:
        super().__init__()
        us = {}
        for nu in range(1, corr + 1):
            u = torch.tensor(so3.u_matrix(tuple(coupling), tuple(ir_out), nu), dtype=torch.get_default_dtype())
            if ir_out[0] == 0:
                u = u[0]  # the reference squeezes the M axis for L = 0 (cg.py:117-131)
            us[str(nu)] = torch.nn.Parameter(u, requires_grad=False)
        self.U_matrices = torch.nn.ParameterDict(us)
        self.weights = torch.nn.ParameterList()
        for nu in range(corr, 0, -1):
            k = us[str(nu)].shape[-1]
            w = torch.nn.Parameter(torch.randn(n_elem, k, n_feat) / k)
            if nu == corr:
                self.weights_max = w
            else:
                self.weights.append(w)
        self.corr = corr

    def by_nu(self):
        """[W_1, ..., W_corr]"""
        out = [None] * self.corr
        out[self.corr - 1] = self.weights_max
        for i, w in enumerate(self.weights):
            out[self.corr - 2 - i] = w
        return out


class SymmetricContraction(torch.nn.Module):
    def __init__(self, irreps_in, irreps_out, correlation: int, num_elements: int) -> None:
        super().__init__()
        irreps_in, irreps_out = _irreps(irreps_in), _irreps(irreps_out)
        coupling = [(l, p) for _, l, p in irreps_in]
        lmax = max(l for l, _ in coupling)
        if coupling != [(l, (-1) ** l) for l in range(lmax + 1)]:
            raise NotImplementedError("symmetric contraction expects coupling irreps 0e+1o+2e+...")
        n_feat = sum(m for m, l, p in irreps_in if (l, p) == (0, 1))
        outs = [(l, p) for _, l, p in irreps_out]
        self.plan = SymPlan(lmax, outs, correlation, n_feat, num_elements)
        self.contractions = torch.nn.ModuleList(
            [Contraction(coupling, o, correlation, n_feat, num_elements) for o in outs]
        )

    def forward(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        ws = self.plan.reduce_weights([w for c in self.contractions for w in c.by_nu()])
        return ops.sym_contract(x, y, ws, self.plan.handle)


class MessageBlock(torch.nn.Module):
    """reference blocks.py:99-121"""

    def __init__(self, interaction_irreps, node_attrs_irreps, hidden_irreps, correlation: int) -> None:
        super().__init__()
        self.symmetric_contractions = SymmetricContraction(
            interaction_irreps, hidden_irreps, correlation, so3.irreps_dim(_irreps(node_attrs_irreps))
        )

    def forward(self, a_i: torch.Tensor, node_attrs: torch.Tensor) -> torch.Tensor:
        return self.symmetric_contractions(a_i, node_attrs)


class NodeUpdateBlock(torch.nn.Module):
    """reference blocks.py:54-96"""

    def __init__(self, node_attrs_irreps, node_feats_irreps, hidden_irreps, residual_connection: bool) -> None:
        super().__init__()
        self.linear = IrrepLinear(hidden_irreps, hidden_irreps)
        self.residual_connection_layer = (
            ElementLinear(node_feats_irreps, so3.irreps_dim(_irreps(node_attrs_irreps)), hidden_irreps)
            if residual_connection else None
        )

    def forward(self, m_i, node_feats, node_attrs, residual=None, species=None):
        """residual: (tensor, event) if the caller already ran ``residual_connection_layer`` on the side stream (it depends
        on the layer's INPUT features only, so MACE.forward starts it before the interaction block)"""
        out = self.linear(m_i)
        if self.residual_connection_layer is not None:
            if residual is None:
                out = out + self.residual_connection_layer(node_feats, node_attrs, species)
            else:
                torch.cuda.current_stream().wait_event(residual[1])
                out = out + residual[0]
        return out


class MACE(torch.nn.Module):
    """reference mace.py:18-175 — same constructor, same ``forward(data) -> data`` contract (writes ``node_feats_{i}``)."""

    def __init__(self, cut_off: float, num_bessel: int, num_polynomial_cutoff: int, max_ell: int, num_interactions: int,
                 num_node_feats: int, num_edge_feats: int, hidden_irreps: str, avg_num_neighbours: float,
                 correlation: int, **kwargs: Any) -> None:
        super().__init__()
        self.r_max = cut_off
        self.max_ell = max_ell
        hidden = _irreps(hidden_irreps)
        C = sum(m for m, l, p in hidden if (l, p) == (0, 1))
        attrs = [(num_node_feats, 0, 1)]
        scalars = [(C, 0, 1)]
        sh = so3.sh_irreps(max_ell)
        edge_feats = [(num_bessel + num_edge_feats, 0, 1)]
        self.node_embedding = IrrepLinear(attrs, scalars)
        self.radial_embedding = RadialEmbeddingBlock(cut_off, num_bessel, num_polynomial_cutoff)
        inter = [(C, l, (-1) ** l) for l in range(max_ell + 1)]
        self.interactions = torch.nn.ModuleList()
        self.messages = torch.nn.ModuleList()
        self.node_updates = torch.nn.ModuleList()
        self.out_irreps = []
        for i in range(num_interactions):
            first = i == 0
            last = (i == num_interactions - 1) and not first
            node_in = scalars if first else hidden
            hid_out = scalars if last else hidden
            self.interactions.append(InteractionBlock(node_in, attrs, sh, edge_feats, inter, avg_num_neighbours, first))
            self.messages.append(MessageBlock(inter, attrs, hid_out, correlation))
            self.node_updates.append(NodeUpdateBlock(attrs, node_in, hid_out, not first))
            self.out_irreps.append(so3.irreps_str(hid_out))
        self._el_plans = None

    def forward(self, data: dict) -> dict:
        ops.reset_pending()
        pos = data["coordinates"]
        edge_index = data["edge_index"]
        attrs = data["node_attrs"]
        cell = shift = None
        if "cell" in data:
            cell = data["cell"][:3].contiguous()  # one cell for the whole batch: only rows 0-2 are read (utils.py:88-90)
            shift = data["cell_shift_vector"].contiguous()
        plan = ops.edge_plan_for(data, pos.shape[0], scatter_row=1)
        if self._el_plans is None:
            self._el_plans = [m.plan for m in self.modules() if isinstance(m, ElementLinear)]
        species = ops.species_plan_for(data, self._el_plans)  # nodes grouped by element: tensor-core element linears
        re = self.radial_embedding
        Y, B = ops.edge_featurise(pos, plan.snd, plan.rcv, shift, cell, re.bessel_fn.bessel_weights, re.r_max,
                                  re.prefactor, 1.0, re.p, self.max_ell)
        if "edge_attrs" in data:
            B = torch.cat([B, data["edge_attrs"]], dim=-1)
        # The radial MLPs of ALL layers depend on the edge features only: they run on a side stream, concurrently with the
        # node-level chain (embedding, per-irrep mixing, contraction), whose GEMMs over ~1e3 rows fill a fraction of the
        # SMs.  Autograd runs the backward of an op on the stream its forward ran on, so the cotangent and
        # double-backward passes get the same two branches; inside a captured step they become parallel graph branches.
        radial = [None] * len(self.interactions)
        if side_stream_ok(B):
            radial = radial_on_side_stream([inter.net_R_channels for inter in self.interactions], B)
        h = self.node_embedding(attrs)
        for i, (inter, msg, upd) in enumerate(zip(self.interactions, self.messages, self.node_updates)):
            res = None
            if radial[i] is not None and upd.residual_connection_layer is not None:
                res = on_side_stream(lambda h_, a_, layer=upd.residual_connection_layer: layer(h_, a_, species), h, attrs)
            a = inter(h, attrs, Y, B, edge_index, edge_plan=plan, radial=radial[i], species=species)
            h = upd(msg(a, attrs), h, attrs, residual=res, species=species)
            data[f"node_feats_{i}"] = h
        return data


class NonLinearReadoutBlock(torch.nn.Module):
    """reference blocks.py:11-28"""

    def __init__(self, irreps_in, MLP_irreps, irreps_out) -> None:
        super().__init__()
        self.linear_1 = IrrepLinear(irreps_in, MLP_irreps)
        self.linear_2 = IrrepLinear(MLP_irreps, irreps_out)

    def forward(self, x):
        return self.linear_2(ops.act(self.linear_1(x), 0, SILU_2MOM))


class ScaleShiftBlock(torch.nn.Module):
    def __init__(self, scale, shift) -> None:
        super().__init__()
        self.register_buffer("scale", torch.tensor(1.0 if scale is None else float(scale)))
        self.register_buffer("shift", torch.tensor(0.0 if shift is None else float(shift)))

    def forward(self, x):
        return self.scale * x + self.shift


class ReadoutHead(torch.nn.Module):
    """reference mace.py:178-214"""

    pooled = False

    def __init__(self, list_in_irreps, mlp_irreps, out_irreps, scaling_std, scaling_mean) -> None:
        super().__init__()
        self.readouts = torch.nn.ModuleList()
        n = len(list_in_irreps)
        for i, irr in enumerate(list_in_irreps):
            if i < n - 1:
                self.readouts.append(IrrepLinear(_irreps(irr), _irreps(out_irreps)))
            else:
                self.readouts.append(NonLinearReadoutBlock(_irreps(irr), _irreps(mlp_irreps), _irreps(out_irreps)))
        self.scale_shift = ScaleShiftBlock(scaling_std, scaling_mean)

    def forward(self, data: dict) -> torch.Tensor:
        x = None
        feats = [data[f"node_feats_{i}"] for i in range(len(self.readouts))]
        small = bool(SIDE_STREAM[0]) and feats[0].is_cuda and 256 <= feats[0].shape[0] <= 16384
        # the per-layer readouts are independent of each other: all but the last run on the side stream (their cotangents
        # then overlap the backward pass of the later layers)
        early = [on_side_stream(ro, f) if small else None for ro, f in zip(self.readouts[:-1], feats[:-1])]
        ys = [self.readouts[-1](feats[-1])]
        for i, ro in enumerate(self.readouts[:-1]):
            if early[i] is None:
                ys.append(ro(feats[i]))
            else:
                torch.cuda.current_stream().wait_event(early[i][1])
                ys.append(early[i][0])
        for y in ys[1:] + ys[:1]:  # layer order, as the reference sums them (mace.py:203-214)
            x = y if x is None else x + y
        x = self.scale_shift(x)
        if self.pooled:
            x = ops.pool_sum(x, data["batch"], int(data["num_graphs"]))
        return x


class PooledReadoutHead(ReadoutHead):
    """reference mace.py:217-260"""

    pooled = True


class PooledMACEModule(torch.nn.Module):
    """The forward slice of the reference's ``PooledMACEModule`` (models/mace/supervised/mace_module.py:13-122):
    backbone + graph-scalar head (+ total_atomic_energy).  Losses / optimiser wiring stay with the caller."""

    def __init__(self, mlp_irreps: str = "16x0e", scaling_mean: float = 0.0, scaling_std: float = 1.0, n_tasks: int = 1,
                 **mace_kwargs: Any) -> None:
        super().__init__()
        self.mace = MACE(**mace_kwargs)
        self.graph_scalars_head = PooledReadoutHead(self.mace.out_irreps, mlp_irreps, f"{n_tasks}x0e", scaling_std, scaling_mean)

    def forward(self, data: dict) -> dict:
        data = self.mace(data)
        e = self.graph_scalars_head(data)
        if "total_atomic_energy" in data:
            e = e + data["total_atomic_energy"].unsqueeze(-1)
        return {"y_graph_scalars": e}


def compute_forces_by_gradient(energy: torch.Tensor, coordinates: torch.Tensor, training=False) -> torch.Tensor:
    """reference lightning/module.py:28-48.  The reference method reads ``self.training``; this free function takes it
    as ``training`` — a bool, or the module itself (its ``.training`` flag is read), so that force-matching training
    (``create_graph``) written against the reference keeps working: ``compute_forces_by_gradient(e, pos, module)``."""
    if isinstance(training, torch.nn.Module):
        training = training.training
    training = bool(training)
    with ops.input_grads_only():  # only d/d coordinates is asked for: no weight gradients in the force pass
        (grad,) = torch.autograd.grad([energy], [coordinates], grad_outputs=[torch.ones_like(energy)],
                                      retain_graph=training, create_graph=training, allow_unused=True)
    if grad is None:
        raise RuntimeWarning("Gradient is None")
    return -1 * grad
