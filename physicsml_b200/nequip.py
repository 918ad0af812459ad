"""Host-side mirror of the reference's NequIP module interface, running on the B200 kernels.

Class names, constructor arguments, the batch-dict contract (``forward(data) -> data`` writing ``node_feats``,
``edge_feats``, ``edge_attrs``) and every state-dict key follow ``physicsml.models.nequip.modules`` (reference
``src/physicsml/models/nequip/modules/nequip.py:14-180``, ``convnet_layer.py:26-112``, ``interaction_block.py:7-108``,
``_gate.py:33-163``, ``radial.py:4-76``), so the reference's fixture checkpoint loads with ``strict=True``.

The convolution shares the fused edge kernels with MACE: K2 (featurise, NequIP radial convention), K3 (radial MLP),
K4 (tensor product + scatter, NequIP's direction: gather ``edge_index[1]``, scatter ``edge_index[0]``), K5a/K5b (linear
mixing, element-indexed self connection); the gate non-linearity is ``csrc/gate.cu``.  No eager fallback.
"""
from __future__ import annotations

import math
from typing import Any

import torch

from . import ops, so3
from .mace import ElementLinear, IrrepLinear, RadialMLP, ScaleShiftBlock, _TPBuffers, _irreps
from .plans import GatePlan, TPPlan


def tp_path_exists(irreps_in1, irreps_in2, ir_out) -> bool:
    """reference convnet_layer.py:9-23"""
    l3, p3 = ir_out
    for _, l1, p1 in irreps_in1:
        for _, l2, p2 in irreps_in2:
            if abs(l1 - l2) <= l3 <= l1 + l2 and p1 * p2 == p3:
                return True
    return False


class BesselBasis(torch.nn.Module):
    """reference nequip/modules/radial.py:4-30 (weights pi*n, argument w*x/r_max, prefactor 2/r_max); evaluated in K2"""

    def __init__(self, r_max: float, num_basis: int = 8, trainable: bool = False) -> None:
        super().__init__()
        w = math.pi * torch.arange(1, num_basis + 1, dtype=torch.get_default_dtype())
        if trainable:
            self.bessel_weights = torch.nn.Parameter(w)
        else:
            self.register_buffer("bessel_weights", w)
        self.r_max = float(r_max)
        self.prefactor = 2.0 / r_max


class RadialEmbeddingBlock(torch.nn.Module):
    def __init__(self, r_max: float, num_bessel: int, num_polynomial_cutoff: int, trainable: bool) -> None:
        super().__init__()
        self.bessel_fn = BesselBasis(r_max, num_bessel, trainable)
        self.r_max, self.p, self.out_dim, self.trainable = float(r_max), int(num_polynomial_cutoff), num_bessel, trainable


class _GateBuffers(torch.nn.Module):
    """state-dict parity with o3.ElementwiseTensorProduct (``equivariant_nonlin.mul``): empty weight + output mask"""

    def __init__(self, dim: int) -> None:
        super().__init__()
        self.register_buffer("weight", torch.zeros(0))
        self.register_buffer("output_mask", torch.ones(dim))


class Gate(torch.nn.Module):
    """reference _gate.py:33-163 with the activations ConvNetLayer chooses (convnet_layer.py:66-77)"""

    def __init__(self, irreps_scalars, irreps_gates, irreps_gated) -> None:
        super().__init__()
        self.plan = GatePlan(irreps_scalars, irreps_gates, irreps_gated)
        self.irreps_in = self.plan.irreps_in
        self.irreps_out = self.plan.irreps_out
        self.mul = _GateBuffers(so3.irreps_dim(_irreps(irreps_gated)))

    def forward(self, features: torch.Tensor) -> torch.Tensor:
        return ops.gate(features, self.plan.handle)


class InteractionBlock(torch.nn.Module):
    """reference interaction_block.py:7-108"""

    def __init__(self, interaction_irreps_in, interaction_irreps_out, node_attrs_irreps, edge_feats_irreps,
                 edge_attrs_irreps, avg_num_neighbours: float | None = None, self_connection: bool = True) -> None:
        super().__init__()
        node = _irreps(interaction_irreps_in)
        out = _irreps(interaction_irreps_out)
        attrs = _irreps(node_attrs_irreps)
        sh = _irreps(edge_attrs_irreps)
        self.avg_num_neighbours = avg_num_neighbours
        lmax = max(l for _, l, _ in sh)
        sh_parity = -1 if any(p == -1 for _, _, p in sh) else 1
        self.linear_1 = IrrepLinear(node, node)
        self.tp_plan = TPPlan(node, lmax, out, sh_parity=sh_parity)
        self.tp = _TPBuffers(self.tp_plan)
        n_edge_feats = so3.irreps_dim(_irreps(edge_feats_irreps))
        self.fc = RadialMLP([n_edge_feats, 64, 64, self.tp_plan.NP * self.tp_plan.C])
        # x.div(avg_num_neighbours ** 0.5) (interaction_block.py:99-100) is folded into linear_2's alpha
        scale = 1.0 if avg_num_neighbours is None else 1.0 / math.sqrt(avg_num_neighbours)
        self.linear_2 = IrrepLinear(so3.simplify_irreps(self.tp_plan.out_irreps), out, scale=scale)
        self.self_connection_layer = ElementLinear(node, so3.irreps_dim(attrs), out) if self_connection else None

    def forward(self, data: dict) -> dict:
        x = data["node_feats"]
        plan = ops.edge_plan_for(data, x.shape[0], scatter_row=0)
        gather, rowptr, perm = plan.view(0)  # gather edge_index[1], scatter to edge_index[0]
        pre = data.get("_radial")
        ready = None
        if pre is not None and id(self) in pre:  # computed on the side stream by Nequip.forward
            weight, ready = pre.pop(id(self))
        else:
            weight = self.fc(data["edge_feats"])
        sc = sc_ready = None
        if self.self_connection_layer is not None:
            if pre is not None:  # small-graph mode: the self connection is independent of the convolution -> side stream
                from .mace import on_side_stream

                sp = data.get("_species")
                sc, sc_ready = on_side_stream(lambda x_, a_: self.self_connection_layer(x_, a_, sp), x, data["node_attrs"])
            else:
                sc = self.self_connection_layer(x, data["node_attrs"], data.get("_species"))
        h = self.linear_1(x)
        if ready is not None:
            torch.cuda.current_stream().wait_event(ready)  # join right before the first consumer of the radial weights
        a = ops.tp_scatter(h, data["edge_attrs"], weight, gather, rowptr, perm, self.tp_plan.handle)
        a = self.linear_2(a)
        if sc is not None:
            if sc_ready is not None:
                torch.cuda.current_stream().wait_event(sc_ready)
            a = a + sc
        data["node_feats"] = a
        return data


class ConvNetLayer(torch.nn.Module):
    """reference convnet_layer.py:26-112"""

    def __init__(self, interaction_irreps_in, node_attrs_irreps, edge_feats_irreps, edge_attrs_irreps, hidden_irreps,
                 avg_num_neighbours: float | None = None, self_connection: bool = True, num_layers: int = 3,
                 resnet: bool = False) -> None:
        super().__init__()
        node = _irreps(interaction_irreps_in)
        sh = _irreps(edge_attrs_irreps)
        hidden = _irreps(hidden_irreps)
        self.hidden_irreps, self.num_layers = hidden, num_layers
        scalars = [(m, l, p) for m, l, p in hidden if l == 0 and tp_path_exists(node, sh, (l, p))]
        gated = [(m, l, p) for m, l, p in hidden if l > 0 and tp_path_exists(node, sh, (l, p))]
        layer_out = so3.simplify_irreps(scalars + gated)
        gate_p = 1 if tp_path_exists(node, sh, (0, 1)) else -1
        gates = [(m, 0, gate_p) for m, _, _ in gated]
        self.equivariant_nonlin = Gate(scalars, gates, gated)
        self.conv_irreps_out = self.equivariant_nonlin.irreps_in
        self.resnet = bool(resnet and layer_out == so3.simplify_irreps(node))
        self.conv = InteractionBlock(node, self.conv_irreps_out, node_attrs_irreps, edge_feats_irreps, sh,
                                     avg_num_neighbours, self_connection)

    def forward(self, data: dict) -> dict:
        old_x = data["node_feats"]
        data = self.conv(data)
        data["node_feats"] = self.equivariant_nonlin(data["node_feats"])
        if self.resnet:
            data["node_feats"] = old_x + data["node_feats"]
        return data


class Nequip(torch.nn.Module):
    """reference nequip.py:14-127 — same constructor, same ``forward(data) -> data`` contract"""

    def __init__(self, cut_off: float, num_layers: int, max_ell: int, parity: bool, num_features: int, num_bessel: int,
                 bessel_basis_trainable: bool, num_polynomial_cutoff: int, self_connection: bool, resnet: bool,
                 num_node_feats: int, num_edge_feats: int, avg_num_neighbours: float, **kwargs: Any) -> None:
        super().__init__()
        self.r_max, self.max_ell = cut_off, max_ell
        attrs = [(num_node_feats, 0, 1)]
        feats = [(num_features, 0, 1)]
        edge_feats = [(num_bessel + num_edge_feats, 0, 1)]
        self.sh_parity = -1 if parity else 1
        sh = so3.sh_irreps(max_ell, self.sh_parity)
        self.node_embedding = IrrepLinear(attrs, feats)
        self.radial_embedding = RadialEmbeddingBlock(cut_off, num_bessel, num_polynomial_cutoff, bessel_basis_trainable)
        hidden = [(num_features, l, 1) for l in range(max_ell + 1)]
        if parity:
            hidden += [(num_features, l, -1) for l in range(max_ell + 1)]
        hidden = so3.simplify_irreps(so3.sort_irreps(hidden)[0])
        node_in = feats
        self.out_irreps = []
        self.conv_layers = torch.nn.ModuleList()
        for _ in range(num_layers):
            layer = ConvNetLayer(node_in, attrs, edge_feats, sh, hidden, avg_num_neighbours, self_connection, num_layers, resnet)
            self.conv_layers.append(layer)
            present = {(l, p) for _, l, p in layer.conv_irreps_out}
            node_in = [(m, l, p) for m, l, p in hidden if (l, p) in present]
            self.out_irreps.append(so3.irreps_str(node_in))
        self._el_plans = None

    def forward(self, data: dict) -> dict:
        ops.reset_pending()
        pos = data["coordinates"]
        cell = shift = None
        if "cell" in data:
            cell = data["cell"][:3].contiguous()
            shift = data["cell_shift_vector"].contiguous()
        # the plan travels to the layers under a private key that lives in `data` for the duration of this call only, so
        # a caller that reuses the dict after rebuilding its neighbour list never sees a stale plan (the reference
        # recomputes everything from edge_index on every call)
        supplied = data.get("_edge_plan")
        plan = ops.edge_plan_for(data, pos.shape[0], scatter_row=0)
        data["_edge_plan"] = plan
        if self._el_plans is None:
            self._el_plans = [m.plan for m in self.modules() if isinstance(m, ElementLinear)]
        data["_species"] = ops.species_plan_for(data, self._el_plans)  # nodes grouped by element (tensor-core self connections)
        re = self.radial_embedding
        w = re.bessel_fn.bessel_weights
        data["node_feats"] = self.node_embedding(data["node_attrs"])
        Y, B = ops.edge_featurise(pos, plan.snd, plan.rcv, shift, cell, w, re.r_max, re.bessel_fn.prefactor,
                                  1.0 / re.r_max, re.p, self.max_ell)
        data["edge_feats"] = torch.cat([B, data["edge_attrs"]], dim=-1) if "edge_attrs" in data else B
        data["edge_attrs"] = Y
        from .mace import radial_on_side_stream, side_stream_ok

        if side_stream_ok(data["edge_feats"]):  # all layers' radial MLPs next to the node-level chain (see mace.MACE.forward)
            convs = [layer.conv for layer in self.conv_layers]
            data["_radial"] = dict(zip((id(c) for c in convs), radial_on_side_stream([c.fc for c in convs], data["edge_feats"])))
        for layer in self.conv_layers:
            data = layer(data)
        data.pop("_radial", None)
        data.pop("_species", None)
        if supplied is None:
            data.pop("_edge_plan", None)
        else:
            data["_edge_plan"] = supplied
        return data


class ReadoutHead(torch.nn.Module):
    """reference nequip.py:130-150"""

    pooled = False

    def __init__(self, irrreps_in, mlp_irreps, out_irreps, scaling_std, scaling_mean) -> None:
        super().__init__()
        self.linear = IrrepLinear(_irreps(irrreps_in), _irreps(mlp_irreps))
        self.output_layer = IrrepLinear(_irreps(mlp_irreps), _irreps(out_irreps))
        self.scale_shift = ScaleShiftBlock(scaling_std, scaling_mean)

    def forward(self, data: dict) -> torch.Tensor:
        x = self.scale_shift(self.output_layer(self.linear(data["node_feats"])))
        if self.pooled:
            x = ops.pool_sum(x, data["batch"], int(data["num_graphs"]))
        return x


class PooledReadoutHead(ReadoutHead):
    """reference nequip.py:153-180"""

    pooled = True


class PooledNequipModule(torch.nn.Module):
    """The forward slice of the reference's ``PooledNequipModule`` (models/nequip/supervised/nequip_module.py:15-128):
    backbone + graph-scalar head (+ total_atomic_energy)."""

    def __init__(self, mlp_irreps: str = "16x0e", scaling_mean: float = 0.0, scaling_std: float = 1.0, n_tasks: int = 1,
                 **nequip_kwargs: Any) -> None:
        super().__init__()
        self.nequip = Nequip(**nequip_kwargs)
        self.graph_scalars_head = PooledReadoutHead(self.nequip.out_irreps[-1], mlp_irreps, f"{n_tasks}x0e", scaling_std, scaling_mean)

    def forward(self, data: dict) -> dict:
        data = self.nequip(data)
        e = self.graph_scalars_head(data)
        if "total_atomic_energy" in data:
            e = e + data["total_atomic_energy"].unsqueeze(-1)
        return {"y_graph_scalars": e}
