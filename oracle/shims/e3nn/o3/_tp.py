"""``o3.Linear`` / ``o3.TensorProduct`` / ``o3.FullyConnectedTensorProduct`` / ``o3.ElementwiseTensorProduct``
restated densely (plain einsum) from the published behaviour of e3nn==0.5.1 — SURVEY.md Appendix A.5-A.7, A.10.

TEST INFRASTRUCTURE ONLY (oracle).  Reference call sites: ``src/physicsml/models/mace/modules/blocks.py:65-78,139-188``,
``mace.py:57-60``, ``nequip/modules/interaction_block.py:22-82``, ``nequip/modules/_gate.py:122``.
Parameter / buffer names mirror e3nn's so the reference's fixture checkpoints load with ``strict=True``.
"""
from __future__ import annotations

import math
from typing import NamedTuple

import torch

from ._irreps import Irreps
from ._wigner import wigner_3j


class Instruction(NamedTuple):
    i_in1: int
    i_in2: int
    i_out: int
    connection_mode: str
    has_weight: bool
    path_weight: float
    path_shape: tuple


class Linear(torch.nn.Module):
    """out[n, i_out, w, m] = sum_paths sum_u W[u, w] x[n, i_in, u, m] / sqrt(fan_in(i_out))."""

    def __init__(self, irreps_in, irreps_out, internal_weights=None, shared_weights=None, **kwargs) -> None:
        super().__init__()
        self.irreps_in = Irreps(irreps_in)
        self.irreps_out = Irreps(irreps_out)
        self.paths = [
            (i_in, i_out)
            for i_in, (_, ir_in) in enumerate(self.irreps_in)
            for i_out, (_, ir_out) in enumerate(self.irreps_out)
            if ir_in == ir_out
        ]
        fan_in = [0] * len(self.irreps_out)
        for i_in, i_out in self.paths:
            fan_in[i_out] += self.irreps_in[i_in].mul
        self.fan_in = fan_in
        self.weight_numel = sum(self.irreps_in[i].mul * self.irreps_out[o].mul for i, o in self.paths)
        self.weight = torch.nn.Parameter(torch.randn(self.weight_numel))
        self.bias = torch.nn.Parameter(torch.zeros(0))
        mask = torch.zeros(self.irreps_out.dim)
        used = {o for _, o in self.paths}
        for o, sl in enumerate(self.irreps_out.slices()):
            if o in used:
                mask[sl] = 1.0
        self.register_buffer("output_mask", mask)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        sl_in, sl_out = self.irreps_in.slices(), self.irreps_out.slices()
        lead = x.shape[:-1]
        outs: list = [None] * len(self.irreps_out)
        off = 0
        for i_in, i_out in self.paths:
            mi, mo = self.irreps_in[i_in], self.irreps_out[i_out]
            w = self.weight[off : off + mi.mul * mo.mul].reshape(mi.mul, mo.mul)
            off += mi.mul * mo.mul
            xi = x[..., sl_in[i_in]].reshape(*lead, mi.mul, mi.ir.dim)
            y = torch.einsum("uw,...ui->...wi", w, xi) * (1.0 / math.sqrt(self.fan_in[i_out]))
            y = y.reshape(*lead, mo.dim)
            outs[i_out] = y if outs[i_out] is None else outs[i_out] + y
        for o, sl in enumerate(sl_out):
            if outs[o] is None:
                outs[o] = x.new_zeros(*lead, sl.stop - sl.start)
        return torch.cat(outs, dim=-1) if outs else x.new_zeros(*lead, 0)


class TensorProduct(torch.nn.Module):
    """Modes 'uvu', 'uvw', 'uuu'; irrep_normalization='component', path_normalization='element'."""

    def __init__(
        self,
        irreps_in1,
        irreps_in2,
        irreps_out,
        instructions,
        in1_var=None,
        in2_var=None,
        out_var=None,
        irrep_normalization: str = "component",
        path_normalization: str = "element",
        internal_weights=None,
        shared_weights=None,
        **kwargs,
    ) -> None:
        super().__init__()
        assert irrep_normalization == "component" and path_normalization == "element"
        self.irreps_in1 = Irreps(irreps_in1)
        self.irreps_in2 = Irreps(irreps_in2)
        self.irreps_out = Irreps(irreps_out)
        if shared_weights is False and internal_weights is None:
            internal_weights = False
        if shared_weights is None and internal_weights is None:
            shared_weights = internal_weights = True
        if shared_weights is None:
            shared_weights = True
        if internal_weights is None:
            internal_weights = shared_weights
        assert shared_weights or not internal_weights
        self.internal_weights = internal_weights
        self.shared_weights = shared_weights

        ins = []
        for x in instructions:
            x = tuple(x)
            if len(x) == 5:
                x = x + (1.0,)
            i1, i2, io, mode, has_w, pw = x
            m1, m2, mo = self.irreps_in1[i1].mul, self.irreps_in2[i2].mul, self.irreps_out[io].mul
            shape = {"uvw": (m1, m2, mo), "uvu": (m1, m2), "uvv": (m1, m2), "uuw": (m1, mo), "uuu": (m1,)}[mode]
            ins.append(Instruction(i1, i2, io, mode, has_w, pw, shape))

        def n_el(i: Instruction) -> int:
            m1, m2 = self.irreps_in1[i.i_in1].mul, self.irreps_in2[i.i_in2].mul
            return {"uvw": m1 * m2, "uvu": m2, "uvv": m1, "uuw": m1, "uuu": 1}[i.connection_mode]

        norm = []
        for i in ins:
            alpha = self.irreps_out[i.i_out].ir.dim
            x = sum(n_el(j) for j in ins if j.i_out == i.i_out)
            if x > 0:
                alpha /= x
            alpha *= i.path_weight
            norm.append(math.sqrt(alpha))
        self.instructions = [
            Instruction(i.i_in1, i.i_in2, i.i_out, i.connection_mode, i.has_weight, c, i.path_shape)
            for i, c in zip(ins, norm)
        ]
        self.weight_numel = sum(math.prod(i.path_shape) for i in self.instructions if i.has_weight)
        if internal_weights and self.weight_numel > 0:
            self.weight = torch.nn.Parameter(torch.randn(self.weight_numel))
        else:
            self.register_buffer("weight", torch.zeros(0))
        mask = torch.zeros(self.irreps_out.dim)
        for i in self.instructions:
            mask[self.irreps_out.slices()[i.i_out]] = 1.0
        self.register_buffer("output_mask", mask)
        # e3nn's generated code keeps the 3j tensors of every path with l1,l2,l3 > 0 as buffers of a submodule
        self._compiled_main_left_right = torch.nn.Module()
        for i in self.instructions:
            l1 = self.irreps_in1[i.i_in1].ir.l
            l2 = self.irreps_in2[i.i_in2].ir.l
            l3 = self.irreps_out[i.i_out].ir.l
            name = f"_w3j_{l1}_{l2}_{l3}"
            if min(l1, l2, l3) > 0 and not hasattr(self._compiled_main_left_right, name):
                self._compiled_main_left_right.register_buffer(name, wigner_3j(l1, l2, l3))

    def forward(self, x1: torch.Tensor, x2: torch.Tensor, weight: torch.Tensor | None = None) -> torch.Tensor:
        if weight is None:
            weight = self.weight
        lead = x1.shape[:-1]
        s1, s2, so = self.irreps_in1.slices(), self.irreps_in2.slices(), self.irreps_out.slices()
        outs: list = [None] * len(self.irreps_out)
        off = 0
        for i in self.instructions:
            a, b, c = self.irreps_in1[i.i_in1], self.irreps_in2[i.i_in2], self.irreps_out[i.i_out]
            xa = x1[..., s1[i.i_in1]].reshape(*lead, a.mul, a.ir.dim)
            xb = x2[..., s2[i.i_in2]].reshape(*lead, b.mul, b.ir.dim)
            w3j = wigner_3j(a.ir.l, b.ir.l, c.ir.l, dtype=x1.dtype, device=x1.device)
            n = math.prod(i.path_shape)
            w = None
            if i.has_weight:
                w = weight[..., off : off + n]
                off += n
                w = w.reshape(*w.shape[:-1], *i.path_shape)
            batched_w = w is not None and w.dim() > len(i.path_shape)
            z = "..." if batched_w else ""
            if i.connection_mode == "uvu":
                y = torch.einsum(f"{z}uv,ijk,...ui,...vj->...uk", w, w3j, xa, xb)
            elif i.connection_mode == "uvw":
                y = torch.einsum(f"{z}uvw,ijk,...ui,...vj->...wk", w, w3j, xa, xb)
            elif i.connection_mode == "uuu":
                if w is None:
                    y = torch.einsum("ijk,...ui,...uj->...uk", w3j, xa, xb)
                else:
                    y = torch.einsum(f"{z}u,ijk,...ui,...uj->...uk", w, w3j, xa, xb)
            else:
                raise NotImplementedError(i.connection_mode)
            y = (i.path_weight * y).reshape(*lead, c.dim)
            outs[i.i_out] = y if outs[i.i_out] is None else outs[i.i_out] + y
        for o, sl in enumerate(so):
            if outs[o] is None:
                outs[o] = x1.new_zeros(*lead, sl.stop - sl.start)
        return torch.cat(outs, dim=-1)


class FullyConnectedTensorProduct(TensorProduct):
    def __init__(self, irreps_in1, irreps_in2, irreps_out, **kwargs) -> None:
        irreps_in1, irreps_in2, irreps_out = Irreps(irreps_in1), Irreps(irreps_in2), Irreps(irreps_out)
        instr = [
            (i1, i2, io, "uvw", True, 1.0)
            for i1, (_, ir1) in enumerate(irreps_in1)
            for i2, (_, ir2) in enumerate(irreps_in2)
            for io, (_, iro) in enumerate(irreps_out)
            if iro in ir1 * ir2
        ]
        super().__init__(irreps_in1, irreps_in2, irreps_out, instr, **kwargs)


class ElementwiseTensorProduct(TensorProduct):
    def __init__(self, irreps_in1, irreps_in2, filter_ir_out=None, **kwargs) -> None:
        irreps_in1, irreps_in2 = Irreps(irreps_in1).simplify(), Irreps(irreps_in2).simplify()
        assert irreps_in1.num_irreps == irreps_in2.num_irreps
        a, b = list(irreps_in1), list(irreps_in2)
        i = 0
        while i < len(a):  # split blocks so that multiplicities line up pairwise
            (m1, ir1), (m2, ir2) = a[i], b[i]
            if m1 < m2:
                b[i] = (m1, ir2)
                b.insert(i + 1, (m2 - m1, ir2))
            if m2 < m1:
                a[i] = (m2, ir1)
                a.insert(i + 1, (m1 - m2, ir1))
            i += 1
        out, instr = [], []
        for i, ((mul, ir1), (mul2, ir2)) in enumerate(zip(a, b)):
            assert mul == mul2
            for ir in ir1 * ir2:
                if filter_ir_out is not None and ir not in filter_ir_out:
                    continue
                instr.append((i, i, len(out), "uuu", False))
                out.append((mul, ir))
        super().__init__(Irreps(a), Irreps(b), Irreps(out), instr, **kwargs)
