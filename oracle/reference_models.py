"""The reference's OWN modules (unmodified source, imported by ``oracle/import_reference.py``) wired into the two
pooled energy models exactly as the reference's Lightning modules wire them, without molflux / lightning:

  * ``models/mace/supervised/mace_module.py:30-70,96-108``      MACE backbone + PooledReadoutHead (+ total_atomic_energy)
  * ``models/nequip/supervised/nequip_module.py:37-80,102-113`` NequIP backbone + PooledReadoutHead

TEST / BASELINE INFRASTRUCTURE ONLY (golden generation, ``bench.py --impl reference``, parity tests).
"""
from __future__ import annotations

import torch

from . import import_reference as ir


def available() -> bool:
    return ir.available()


def _o3():
    ir.install_shims()
    from e3nn import o3

    return o3


class ReferencePooledMACE(torch.nn.Module):
    def __init__(self, mlp_irreps: str = "16x0e", **kw) -> None:
        super().__init__()
        mod, o3 = ir.load("physicsml.models.mace.modules.mace"), _o3()
        self.mace = mod.MACE(**kw)
        self.graph_scalars_head = mod.PooledReadoutHead(
            list_in_irreps=self.mace.out_irreps, mlp_irreps=o3.Irreps(mlp_irreps), out_irreps=o3.Irreps("1x0e"),
            scaling_std=1.0, scaling_mean=0.0)

    def forward(self, data):
        e = self.graph_scalars_head(self.mace(data))
        if "total_atomic_energy" in data:
            e = e + data["total_atomic_energy"].unsqueeze(-1)
        return {"y_graph_scalars": e}


class ReferencePooledNequip(torch.nn.Module):
    def __init__(self, mlp_irreps: str = "16x0e", **kw) -> None:
        super().__init__()
        mod, o3 = ir.load("physicsml.models.nequip.modules.nequip"), _o3()
        self.nequip = mod.Nequip(**kw)
        self.graph_scalars_head = mod.PooledReadoutHead(
            irrreps_in=self.nequip.out_irreps[-1], mlp_irreps=o3.Irreps(mlp_irreps), out_irreps=o3.Irreps("1x0e"),
            scaling_std=1.0, scaling_mean=0.0)

    def forward(self, data):
        e = self.graph_scalars_head(self.nequip(data))
        if "total_atomic_energy" in data:
            e = e + data["total_atomic_energy"].unsqueeze(-1)
        return {"y_graph_scalars": e}
