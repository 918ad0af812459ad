"""CPU restatement (plain PyTorch eager, dense einsums) of the reference's MACE energy path.

TEST INFRASTRUCTURE ONLY (oracle): only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product (``physicsml_b200``) never does.

What is restated (reference paths relative to /root/reference/src/physicsml):
  * edge vectors / lengths                    models/utils.py:71-95
  * spherical harmonics + Bessel x cutoff     models/mace/modules/mace.py:141-152, blocks.py:31-51, radial.py:6-49
  * InteractionBlock                          models/mace/modules/blocks.py:124-237, irreps_tools.py:14-67
  * SymmetricContraction / Contraction        models/mace/modules/symmetric_contraction.py:24-243
  * NodeUpdateBlock                           models/mace/modules/blocks.py:54-96
  * readout heads + pooling                   models/mace/modules/mace.py:178-260, blocks.py:11-28,240-256
  * forces = -dE/dx by autograd               lightning/module.py:28-48

Module / parameter names mirror the reference so that one state dict drives the reference (imported over the shims
in the build container), this oracle and the CUDA product.  The e3nn arithmetic underneath is the oracle's own
restatement (``oracle/shims/e3nn``), pinned as described in DESIGN.md ("parity pins").

Validated in the build container against the reference's unmodified ``mace.py`` (tests/golden/make_golden.py writes
the outputs of that run to tests/golden/*.pt; tests/test_oracle_golden.py replays them anywhere).
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

from .ref_cg import u_matrix  # noqa: E402


# --------------------------------------------------------------------------------------------------------------------
# geometry + featurisation
# --------------------------------------------------------------------------------------------------------------------
def edge_vectors(pos, edge_index, cell=None, shift=None):
    """models/utils.py:71-95 — r = pos[receiver] - pos[sender] + shift @ cell[0:3]; |r| clamped at 1e-8."""
    snd, rcv = edge_index[0], edge_index[1]
    r = pos[rcv] - pos[snd]
    if cell is not None and shift is not None:
        r = r + shift[:, 0:1] * cell[0][None] + shift[:, 1:2] * cell[1][None] + shift[:, 2:3] * cell[2][None]
    return torch.norm(r, dim=-1).unsqueeze(-1).clamp(min=1e-8), r


class _Bessel(torch.nn.Module):
    def __init__(self, r_max: float, n: int) -> None:
        super().__init__()
        self.register_buffer("bessel_weights", math.pi * torch.arange(1, n + 1) / r_max)
        self.pref = math.sqrt(2.0 / r_max)

    def forward(self, d):
        return self.pref * (torch.sin(self.bessel_weights * d) / d)


class _Radial(torch.nn.Module):
    """blocks.py:31-51 + radial.py:6-49 (MACE flavour: w_n = n pi / r_max, prefactor sqrt(2/r_max))."""

    def __init__(self, r_max: float, n: int, p: int) -> None:
        super().__init__()
        self.bessel_fn = _Bessel(r_max, n)
        self.r_max, self.p = r_max, p

    def forward(self, d):
        p, x = self.p, d / self.r_max
        env = (
            1.0
            - ((p + 1.0) * (p + 2.0) / 2.0) * torch.pow(x, p)
            + p * (p + 2.0) * torch.pow(x, p + 1)
            - (p * (p + 1.0) / 2) * torch.pow(x, p + 2)
        )
        return self.bessel_fn(d) * (env * (d < self.r_max))


# --------------------------------------------------------------------------------------------------------------------
# interaction
# --------------------------------------------------------------------------------------------------------------------
def uvu_instructions(irreps_node, irreps_sh, target):
    """irreps_tools.py:14-43 — one 'uvu' path per (node irrep, sh irrep, l3 in target); outputs sorted stably."""
    blocks, ins = [], []
    for i, (mul, ir1) in enumerate(irreps_node):
        for j, (_, ir2) in enumerate(irreps_sh):
            for ir3 in ir1 * ir2:
                if ir3 in target:
                    ins.append((i, j, len(blocks)))
                    blocks.append((mul, ir3))
    srt = o3.Irreps(blocks).sort()
    return srt.irreps, [(i, j, srt.p[k], "uvu", True) for i, j, k in ins]


def to_channel_major(x, irreps):
    """irreps_tools.py:47-67 — [N, sum_l C(2l+1)] (irreps layout) -> [N, C, sum_l (2l+1)]."""
    parts, off = [], 0
    for mul, ir in irreps:
        parts.append(x[:, off : off + mul * ir.dim].reshape(x.shape[0], mul, ir.dim))
        off += mul * ir.dim
    return torch.cat(parts, dim=-1)


class _Interaction(torch.nn.Module):
    def __init__(self, node_irreps, attr_irreps, sh_irreps, n_radial, inter_irreps, avg_nb, mix) -> None:
        super().__init__()
        self.avg = avg_nb
        self.inter_irreps = inter_irreps
        self.linear_node_feats = o3.Linear(node_irreps, node_irreps)
        tp_irreps, ins = uvu_instructions(node_irreps, sh_irreps, inter_irreps)
        self.net_R_channels = e3nn_nn.FullyConnectedNet([n_radial, 64, 64, 64, tp_irreps.num_irreps], torch.nn.SiLU())
        self.conv_tp = o3.TensorProduct(
            node_irreps, sh_irreps, tp_irreps, instructions=ins, shared_weights=False, internal_weights=False
        )
        self.linear = o3.Linear(tp_irreps.simplify(), inter_irreps)
        self.mix_layer = o3.FullyConnectedTensorProduct(inter_irreps, attr_irreps, inter_irreps) if mix else None

    def forward(self, h, attrs, Y, B, edge_index):
        snd, rcv = edge_index[0], edge_index[1]
        wh = self.linear_node_feats(h)
        R = self.net_R_channels(B)
        m = self.conv_tp(wh[snd], Y, R)  # [E, d_tp]
        A = torch.zeros(h.shape[0], m.shape[1], dtype=m.dtype, device=m.device).index_add_(0, rcv, m)
        A = self.linear(A) / self.avg
        if self.mix_layer is not None:
            A = self.mix_layer(A, attrs)
        return to_channel_major(A, self.inter_irreps)


# --------------------------------------------------------------------------------------------------------------------
# product basis
# --------------------------------------------------------------------------------------------------------------------
_AX = "wxvnzrtyuops"


class _Contraction(torch.nn.Module):
    """symmetric_contraction.py:81-243 for one output irrep."""

    def __init__(self, coupling, ir_out, corr: int, n_feat: int, n_elem: int) -> None:
        super().__init__()
        self.corr, self.has_m = corr, ir_out.l > 0
        self.U_matrices = torch.nn.ParameterDict(
            {str(nu): torch.nn.Parameter(u_matrix(coupling, ir_out, nu), requires_grad=False) for nu in range(1, corr + 1)}
        )
        self.weights = torch.nn.ParameterList()
        for nu in range(corr, 0, -1):
            k = self.U_matrices[str(nu)].shape[-1]
            w = torch.nn.Parameter(torch.randn(n_elem, k, n_feat) / k)
            if nu == corr:
                self.weights_max = w
            else:
                self.weights.append(w)

    def forward(self, x, y):
        corr, lead = self.corr, int(self.has_m)
        a = _AX[: corr + lead - 1]
        out = torch.einsum(f"{a}ik,ekc,bci,be->bc{a}", self.U_matrices[str(corr)], self.weights_max, x, y)
        for j, w in enumerate(self.weights):
            nu = corr - j - 1
            a = _AX[: nu + lead]
            c = torch.einsum(f"{a}k,ekc,be->bc{a}", self.U_matrices[str(nu)], w, y) + out
            a = _AX[: nu - 1 + lead]
            out = torch.einsum(f"bc{a}i,bci->bc{a}", c, x)
        return out.reshape(out.shape[0], -1)


class _SymContraction(torch.nn.Module):
    def __init__(self, irreps_in, irreps_out, corr: int, n_elem: int) -> None:
        super().__init__()
        coupling = [ir for _, ir in irreps_in]
        n_feat = irreps_in.count((0, 1))
        self.contractions = torch.nn.ModuleList(
            [_Contraction(coupling, ir, corr, n_feat, n_elem) for _, ir in irreps_out]
        )

    def forward(self, x, y):
        return torch.cat([c(x, y) for c in self.contractions], dim=-1)


class _Message(torch.nn.Module):
    def __init__(self, inter_irreps, attr_irreps, hidden, corr) -> None:
        super().__init__()
        self.symmetric_contractions = _SymContraction(inter_irreps, hidden, corr, attr_irreps.dim)

    def forward(self, a, attrs):
        return self.symmetric_contractions(a, attrs)


class _NodeUpdate(torch.nn.Module):
    def __init__(self, attr_irreps, node_irreps, hidden, residual: bool) -> None:
        super().__init__()
        self.linear = o3.Linear(hidden, hidden)
        self.residual_connection_layer = (
            o3.FullyConnectedTensorProduct(node_irreps, attr_irreps, hidden) if residual else None
        )

    def forward(self, m, h, attrs):
        out = self.linear(m)
        if self.residual_connection_layer is not None:
            out = out + self.residual_connection_layer(h, attrs)
        return out


# --------------------------------------------------------------------------------------------------------------------
# model
# --------------------------------------------------------------------------------------------------------------------
class RefMACE(torch.nn.Module):
    """Same constructor arguments and state-dict keys as ``physicsml.models.mace.modules.mace.MACE`` (mace.py:18-131)."""

    def __init__(
        self,
        cut_off: float,
        num_bessel: int,
        num_polynomial_cutoff: int,
        max_ell: int,
        num_interactions: int,
        num_node_feats: int,
        num_edge_feats: int,
        hidden_irreps: str,
        avg_num_neighbours: float,
        correlation: int,
        **_,
    ) -> None:
        super().__init__()
        hidden = o3.Irreps(hidden_irreps)
        C = hidden.count(o3.Irrep(0, 1))
        attr_irreps = o3.Irreps(f"{num_node_feats}x0e")
        scalar_irreps = o3.Irreps(f"{C}x0e")
        self.spherical_harmonics = o3.SphericalHarmonics(
            o3.Irreps.spherical_harmonics(max_ell), normalize=True, normalization="component"
        )
        sh_irreps = self.spherical_harmonics.irreps_out
        self.node_embedding = o3.Linear(attr_irreps, scalar_irreps)
        self.radial_embedding = _Radial(cut_off, num_bessel, num_polynomial_cutoff)
        inter = (sh_irreps * C).sort().irreps.simplify()
        self.interactions = torch.nn.ModuleList()
        self.messages = torch.nn.ModuleList()
        self.node_updates = torch.nn.ModuleList()
        self.out_irreps = []
        for i in range(num_interactions):
            first = i == 0
            last = (i == num_interactions - 1) and not first
            node_in = scalar_irreps if first else hidden
            hid_out = scalar_irreps if last else hidden
            self.interactions.append(
                _Interaction(node_in, attr_irreps, sh_irreps, num_bessel + num_edge_feats, inter, avg_num_neighbours, first)
            )
            self.messages.append(_Message(inter, attr_irreps, hid_out, correlation))
            self.node_updates.append(_NodeUpdate(attr_irreps, node_in, hid_out, not first))
            self.out_irreps.append(hid_out)

    def forward(self, data: dict) -> dict:
        cell = data.get("cell")
        shift = data.get("cell_shift_vector") if cell is not None else None
        d, r = edge_vectors(data["coordinates"], data["edge_index"], cell, shift)
        attrs = data["node_attrs"]
        h = self.node_embedding(attrs)
        Y = self.spherical_harmonics(r)
        B = self.radial_embedding(d)
        if "edge_attrs" in data:
            B = torch.cat([B, data["edge_attrs"]], dim=-1)
        for i, (inter, msg, upd) in enumerate(zip(self.interactions, self.messages, self.node_updates)):
            a = inter(h, attrs, Y, B, data["edge_index"])
            h = upd(msg(a, attrs), h, attrs)
            data[f"node_feats_{i}"] = h
        return data


class _NonLinearReadout(torch.nn.Module):
    """blocks.py:11-28 — Linear -> normalize2mom(SiLU) on scalars -> Linear."""

    def __init__(self, irreps_in, mlp_irreps, irreps_out) -> None:
        super().__init__()
        self.linear_1 = o3.Linear(irreps_in, mlp_irreps)
        self.act = normalize2mom(torch.nn.SiLU())
        self.linear_2 = o3.Linear(mlp_irreps, irreps_out)

    def forward(self, x):
        return self.linear_2(self.act(self.linear_1(x)))


class RefPooledReadoutHead(torch.nn.Module):
    """mace.py:217-260 (+ ``pooled=False`` gives ReadoutHead, mace.py:178-214)."""

    def __init__(self, list_in_irreps, mlp_irreps, out_irreps, scaling_std, scaling_mean, pooled: bool = True) -> None:
        super().__init__()
        self.pooled = pooled
        self.readouts = torch.nn.ModuleList()
        for i, irr in enumerate(list_in_irreps):
            if i < len(list_in_irreps) - 1:
                self.readouts.append(o3.Linear(irr, o3.Irreps(out_irreps)))
            else:
                self.readouts.append(_NonLinearReadout(irr, o3.Irreps(mlp_irreps), o3.Irreps(out_irreps)))
        self.scale_shift = torch.nn.Module()
        self.scale_shift.register_buffer("scale", torch.tensor(1.0 if scaling_std is None else scaling_std))
        self.scale_shift.register_buffer("shift", torch.tensor(0.0 if scaling_mean is None else scaling_mean))

    def forward(self, data):
        x = sum(ro(data[f"node_feats_{i}"]) for i, ro in enumerate(self.readouts))
        x = self.scale_shift.scale * x + self.scale_shift.shift
        if not self.pooled:
            return x
        out = x.new_zeros(int(data["num_graphs"]), x.shape[1])
        return out.index_add_(0, data["batch"], x)


class RefPooledMACE(torch.nn.Module):
    """The energy slice of ``PooledMACEModule`` (models/mace/supervised/mace_module.py:13-122): backbone + graph head."""

    def __init__(self, mlp_irreps="16x0e", scaling_mean=0.0, scaling_std=1.0, n_tasks=1, **mace_kwargs) -> None:
        super().__init__()
        self.mace = RefMACE(**mace_kwargs)
        self.graph_scalars_head = RefPooledReadoutHead(
            self.mace.out_irreps, mlp_irreps, f"{n_tasks}x0e", scaling_std, scaling_mean
        )

    def forward(self, data: dict) -> dict:
        data = self.mace(data)
        e = self.graph_scalars_head(data)
        if "total_atomic_energy" in data:
            e = e + data["total_atomic_energy"].unsqueeze(-1)
        return {"y_graph_scalars": e}


def energy_and_forces(model, data: dict, training: bool = False):
    """lightning/module.py:28-48,130-174 — forces = -dE/dx with create_graph only in training."""
    data = dict(data)
    pos = data["coordinates"].detach().clone().requires_grad_(True)
    data["coordinates"] = pos
    e = model(data)["y_graph_scalars"]
    (g,) = torch.autograd.grad([e], [pos], grad_outputs=[torch.ones_like(e)], retain_graph=training, create_graph=training)
    return e, -g
