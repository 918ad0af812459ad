"""``e3nn.nn.FullyConnectedNet`` and ``e3nn.nn.Extract`` restated.
TEST INFRASTRUCTURE ONLY (oracle).  Call sites: ``mace/modules/blocks.py:157-160``, ``nequip/modules/interaction_block.py:61-64``,
``nequip/modules/_gate.py:26``."""
import math

import torch

from ..math import normalize2mom
from ..o3 import Irreps


class _Layer(torch.nn.Module):
    def __init__(self, h_in: int, h_out: int, act) -> None:
        super().__init__()
        self.weight = torch.nn.Parameter(torch.randn(h_in, h_out))
        self.act = act
        self.h_in = h_in

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = x @ (self.weight / math.sqrt(self.h_in))
        return x if self.act is None else self.act(x)


class FullyConnectedNet(torch.nn.Sequential):
    def __init__(self, hs, act=None, variance_in=1, variance_out=1, out_act=False) -> None:
        super().__init__()
        self.hs = list(hs)
        act = None if act is None else normalize2mom(act)
        n = len(self.hs) - 1
        for i, (h1, h2) in enumerate(zip(self.hs, self.hs[1:])):
            last = i == n - 1
            setattr(self, f"layer{i}", _Layer(h1, h2, act if (not last or out_act) else None))


class Extract(torch.nn.Module):
    def __init__(self, irreps_in, irreps_outs, instructions, squeeze_out: bool = False) -> None:
        super().__init__()
        self.irreps_in = Irreps(irreps_in)
        self.irreps_outs = tuple(Irreps(i) for i in irreps_outs)
        self.instructions = instructions
        self.squeeze_out = squeeze_out
        sl = self.irreps_in.slices()
        self._slices = [[sl[i] for i in ins] for ins in instructions]

    def forward(self, x: torch.Tensor):
        outs = []
        for sls, irr in zip(self._slices, self.irreps_outs):
            if sls:
                outs.append(torch.cat([x[..., s] for s in sls], dim=-1))
            else:
                outs.append(x.new_zeros(*x.shape[:-1], 0))
        if self.squeeze_out and len(outs) == 1:
            return outs[0]
        return tuple(outs)
