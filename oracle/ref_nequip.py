"""CPU restatement (plain PyTorch eager) of the reference's NequIP energy path.

TEST INFRASTRUCTURE ONLY (oracle): only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product (``physicsml_b200``) never does.

What is restated (reference paths relative to /root/reference/src/physicsml/models/nequip):
  * Bessel basis x polynomial cutoff (NequIP flavour)   modules/radial.py:4-76
  * InteractionBlock                                     modules/interaction_block.py:7-108
  * Gate (sort-cut, activations, elementwise product)    modules/_gate.py:10-152, mace/modules/_activation.py:6-117
  * ConvNetLayer (gate construction, resnet)             modules/convnet_layer.py:9-112
  * Nequip backbone                                      modules/nequip.py:14-127
  * PooledReadoutHead                                    modules/nequip.py:153-180, modules/scale_shift.py

Module / parameter names mirror the reference so that one state dict drives the reference (imported over the shims in
the build container), this oracle and the CUDA product.  Validated in the build container against the reference's
unmodified ``nequip.py`` (tests/golden/make_golden.py -> tests/golden/nequip_*.pt; tests/test_oracle_golden.py).
The e3nn arithmetic underneath is the oracle's own restatement (``oracle/shims/e3nn``): parity unpinned at the e3nn
boundary exactly as for MACE.
"""
from __future__ import annotations

import math
import os
import sys

import torch

_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")
if _SHIMS not in sys.path:
    sys.path.insert(0, _SHIMS)

from e3nn import nn as e3nn_nn  # noqa: E402
from e3nn import o3  # noqa: E402
from e3nn.math import normalize2mom  # noqa: E402

from .ref_model import edge_vectors  # noqa: E402


class _Bessel(torch.nn.Module):
    """radial.py:4-30: w_n = n pi, argument w x / r_max, prefactor 2 / r_max (no sqrt, as in the original NequIP)"""

    def __init__(self, r_max: float, n: int, trainable: bool) -> None:
        super().__init__()
        w = math.pi * torch.arange(1, n + 1)
        if trainable:
            self.bessel_weights = torch.nn.Parameter(w)
        else:
            self.register_buffer("bessel_weights", w)
        self.r_max, self.prefactor = r_max, 2.0 / r_max

    def forward(self, x):
        return self.prefactor * (torch.sin(self.bessel_weights * x / self.r_max) / x)


class _Radial(torch.nn.Module):
    def __init__(self, r_max: float, n: int, p: int, trainable: bool) -> None:
        super().__init__()
        self.bessel_fn = _Bessel(r_max, n, trainable)
        self.r_max, self.p = r_max, p

    def forward(self, d):
        p, x = self.p, d / self.r_max
        env = (1.0 - ((p + 1.0) * (p + 2.0) / 2.0) * torch.pow(x, p) + p * (p + 2.0) * torch.pow(x, p + 1)
               - (p * (p + 1.0) / 2) * torch.pow(x, p + 2))
        return self.bessel_fn(d) * (env * (d < self.r_max))


def _path_exists(irreps1, irreps2, ir_out) -> bool:
    ir_out = o3.Irrep(ir_out)
    return any(ir_out in ir1 * ir2 for _, ir1 in o3.Irreps(irreps1).simplify() for _, ir2 in o3.Irreps(irreps2).simplify())


class _Interaction(torch.nn.Module):
    """interaction_block.py:7-108"""

    def __init__(self, irreps_in, irreps_out, attrs, edge_feats, sh, avg, self_connection) -> None:
        super().__init__()
        self.avg = avg
        self.linear_1 = o3.Linear(irreps_in, irreps_in)
        mid, ins = [], []
        for i, (mul, ir_in) in enumerate(irreps_in):
            for j, (_, ir_edge) in enumerate(sh):
                for ir_out in ir_in * ir_edge:
                    if ir_out in irreps_out:
                        ins.append((i, j, len(mid)))
                        mid.append((mul, ir_out))
        srt = o3.Irreps(mid).sort()
        self.tp = o3.TensorProduct(irreps_in, sh, srt.irreps, [(i, j, srt.p[k], "uvu", True) for i, j, k in ins],
                                   shared_weights=False, internal_weights=False)
        self.fc = e3nn_nn.FullyConnectedNet([edge_feats.num_irreps, 64, 64, self.tp.weight_numel], torch.nn.SiLU())
        self.linear_2 = o3.Linear(srt.irreps.simplify(), irreps_out)
        self.self_connection_layer = o3.FullyConnectedTensorProduct(irreps_in, attrs, irreps_out) if self_connection else None

    def forward(self, data):
        w = self.fc(data["edge_feats"])
        x = data["node_feats"]
        src, dst = data["edge_index"][1], data["edge_index"][0]
        sc = self.self_connection_layer(x, data["node_attrs"]) if self.self_connection_layer is not None else None
        x = self.linear_1(x)
        m = self.tp(x[src], data["edge_attrs"], w)
        x = torch.zeros(x.shape[0], m.shape[1], dtype=m.dtype).index_add_(0, dst, m)
        if self.avg is not None:
            x = x.div(self.avg**0.5)
        x = self.linear_2(x)
        if sc is not None:
            x = x + sc
        data["node_feats"] = x
        return data


class _Gate(torch.nn.Module):
    """_gate.py:88-152 with SiLU on even / tanh on odd scalars and gates, each through normalize2mom (_activation.py:40)"""

    def __init__(self, scalars, gates, gated) -> None:
        super().__init__()
        scalars, gates, gated = o3.Irreps(scalars), o3.Irreps(gates), o3.Irreps(gated)
        outs = tuple(o3.Irreps(i).simplify() for i in (scalars, gates, gated))
        cat = sum(outs, o3.Irreps([]))
        groups, i = [], 0
        for o in outs:
            groups.append(tuple(range(i, i + len(o))))
            i += len(o)
        srt = cat.sort()
        self.cut = e3nn_nn.Extract(srt.irreps, outs, [tuple(srt.p[k] for k in g) for g in groups])
        self.irreps_in = srt.irreps.simplify()
        self.irreps_scalars, self.irreps_gates, self.irreps_gated = scalars, gates, gated
        act = lambda ir: normalize2mom(torch.tanh if ir.p == -1 else torch.nn.SiLU())  # noqa: E731
        self.act_s = torch.nn.ModuleList([act(ir) for _, ir in scalars])
        self.act_g = torch.nn.ModuleList([act(ir) for _, ir in gates])
        self.mul = o3.ElementwiseTensorProduct(gated, gates)
        self.irreps_out = scalars + self.mul.irreps_out

    @staticmethod
    def _activate(x, irreps, acts):
        out, off = [], 0
        for (mul, _), a in zip(irreps, acts):
            out.append(a(x[..., off : off + mul]))
            off += mul
        return torch.cat(out, dim=-1) if out else x

    def forward(self, x):
        s, g, y = self.cut(x)
        s = self._activate(s, self.irreps_scalars, self.act_s)
        if g.shape[-1]:
            g = self._activate(g, self.irreps_gates, self.act_g)
            return torch.cat([s, self.mul(y, g)], dim=-1)
        return s


class _ConvNetLayer(torch.nn.Module):
    """convnet_layer.py:26-112"""

    def __init__(self, irreps_in, attrs, edge_feats, sh, hidden, avg, self_connection, resnet) -> None:
        super().__init__()
        scalars = o3.Irreps([(m, ir) for m, ir in hidden if ir.l == 0 and _path_exists(irreps_in, sh, ir)])
        gated = o3.Irreps([(m, ir) for m, ir in hidden if ir.l > 0 and _path_exists(irreps_in, sh, ir)])
        layer_out = (scalars + gated).simplify()
        ir = "0e" if _path_exists(irreps_in, sh, "0e") else "0o"
        gates = o3.Irreps([(m, ir) for m, _ in gated])
        self.equivariant_nonlin = _Gate(scalars, gates, gated)
        self.conv_irreps_out = self.equivariant_nonlin.irreps_in.simplify()
        self.resnet = bool(resnet and layer_out == irreps_in)
        self.conv = _Interaction(irreps_in, self.conv_irreps_out, attrs, edge_feats, sh, avg, self_connection)

    def forward(self, data):
        old = data["node_feats"]
        data = self.conv(data)
        data["node_feats"] = self.equivariant_nonlin(data["node_feats"])
        if self.resnet:
            data["node_feats"] = old + data["node_feats"]
        return data


class RefNequip(torch.nn.Module):
    """nequip.py:14-127"""

    def __init__(self, cut_off, num_layers, max_ell, parity, num_features, num_bessel, bessel_basis_trainable,
                 num_polynomial_cutoff, self_connection, resnet, num_node_feats, num_edge_feats, avg_num_neighbours, **kw):
        super().__init__()
        attrs = o3.Irreps(f"{num_node_feats}x0e")
        feats = o3.Irreps(f"{num_features}x0e")
        edge_feats = o3.Irreps(f"{num_bessel + num_edge_feats}x0e")
        self.spherical_harmonics = o3.SphericalHarmonics(o3.Irreps.spherical_harmonics(max_ell, p=-1 if parity else 1),
                                                         normalize=True, normalization="component")
        sh = self.spherical_harmonics.irreps_out
        self.node_embedding = o3.Linear(attrs, feats)
        self.radial_embedding = _Radial(cut_off, num_bessel, num_polynomial_cutoff, bessel_basis_trainable)
        hid = [(num_features, f"{l}e") for l in range(max_ell + 1)]
        if parity:
            hid += [(num_features, f"{l}o") for l in range(max_ell + 1)]
        hidden = o3.Irreps(hid).sort().irreps.simplify()
        irreps_in = feats
        self.out_irreps = []
        self.conv_layers = torch.nn.ModuleList()
        for _ in range(num_layers):
            layer = _ConvNetLayer(irreps_in, attrs, edge_feats, sh, hidden, avg_num_neighbours, self_connection, resnet)
            self.conv_layers.append(layer)
            irreps_in = o3.Irreps([(m, ir) for m, ir in hidden if ir in layer.conv_irreps_out])
            self.out_irreps.append(irreps_in)

    def forward(self, data):
        data["node_feats"] = self.node_embedding(data["node_attrs"])
        d, r = edge_vectors(data["coordinates"], data["edge_index"], data.get("cell"), data.get("cell_shift_vector"))
        ef = self.radial_embedding(d)
        data["edge_feats"] = torch.cat([ef, data["edge_attrs"]], dim=-1) if "edge_attrs" in data else ef
        data["edge_attrs"] = self.spherical_harmonics(r)
        for layer in self.conv_layers:
            data = layer(data)
        return data


class _Head(torch.nn.Module):
    """nequip.py:153-180 + scale_shift.py"""

    def __init__(self, irreps_in, mlp, out, std, mean) -> None:
        super().__init__()
        self.linear = o3.Linear(irreps_in, mlp)
        self.output_layer = o3.Linear(mlp, out)
        self.scale_shift = torch.nn.Module()
        self.scale_shift.register_buffer("scale", torch.tensor(float(std)))
        self.scale_shift.register_buffer("shift", torch.tensor(float(mean)))

    def forward(self, data):
        x = self.output_layer(self.linear(data["node_feats"]))
        x = self.scale_shift.scale * x + self.scale_shift.shift
        G = int(data["num_graphs"])
        return torch.zeros(G, x.shape[1], dtype=x.dtype).index_add_(0, data["batch"], x)


class RefPooledNequip(torch.nn.Module):
    """nequip_module.py:15-128 (graph-scalar head only)"""

    def __init__(self, mlp_irreps="16x0e", scaling_mean=0.0, scaling_std=1.0, **kw) -> None:
        super().__init__()
        self.nequip = RefNequip(**kw)
        self.graph_scalars_head = _Head(self.nequip.out_irreps[-1], o3.Irreps(mlp_irreps), o3.Irreps("1x0e"), scaling_std, scaling_mean)

    def forward(self, data):
        data = self.nequip(dict(data))
        e = self.graph_scalars_head(data)
        if "total_atomic_energy" in data:
            e = e + data["total_atomic_energy"].unsqueeze(-1)
        return {"y_graph_scalars": e}
