"""Real spherical harmonics up to l=3 in e3nn's basis and ordering.

TEST INFRASTRUCTURE ONLY (oracle).  Restates ``e3nn.o3.SphericalHarmonics`` (e3nn==0.5.1) as used at
``src/physicsml/models/mace/modules/mace.py:50-54,149`` and ``nequip/modules/nequip.py:41-45,122``.
The basis is cross-checked against the pinned Wigner-3j tensors in ``tests/test_oracle_pins.py``
(Y_l is a positive multiple of w3j(1,l-1,l) . Y_1 (x) Y_{l-1}; |Y_l|^2 = 2l+1).
"""
from __future__ import annotations

import math

import torch

from ._irreps import Irreps


def _raw_sh(lmax: int, x: torch.Tensor, y: torch.Tensor, z: torch.Tensor) -> list[torch.Tensor]:
    """'norm'-normalised harmonics (|Y_l| = 1 on the unit sphere), list of [..., 2l+1]."""
    out = [torch.ones_like(x).unsqueeze(-1)]
    if lmax == 0:
        return out
    out.append(torch.stack([x, y, z], dim=-1))
    if lmax == 1:
        return out
    x2, y2, z2 = x * x, y * y, z * z
    x2z2 = x2 + z2
    s3 = math.sqrt(3.0)
    sh20 = s3 * x * z
    sh21 = s3 * x * y
    sh22 = y2 - 0.5 * x2z2
    sh23 = s3 * y * z
    sh24 = (s3 / 2.0) * (z2 - x2)
    out.append(torch.stack([sh20, sh21, sh22, sh23, sh24], dim=-1))
    if lmax == 2:
        return out
    c30 = math.sqrt(30.0) / 6.0
    s5 = math.sqrt(5.0)
    c6 = math.sqrt(6.0) / 4.0
    sh30 = c30 * (sh20 * z + sh24 * x)
    sh31 = s5 * sh20 * y
    sh32 = c6 * (4.0 * y2 - x2z2) * x
    sh33 = 0.5 * y * (2.0 * y2 - 3.0 * x2z2)
    sh34 = c6 * z * (4.0 * y2 - x2z2)
    sh35 = s5 * sh24 * y
    sh36 = c30 * (sh24 * z - sh20 * x)
    out.append(torch.stack([sh30, sh31, sh32, sh33, sh34, sh35, sh36], dim=-1))
    if lmax > 3:
        raise NotImplementedError("oracle spherical harmonics are restated up to l=3 only")
    return out


class SphericalHarmonics(torch.nn.Module):
    def __init__(self, irreps_out, normalize: bool, normalization: str = "integral", irreps_in=None) -> None:
        super().__init__()
        if isinstance(irreps_out, int):
            irreps_out = Irreps([(1, (irreps_out, (-1) ** irreps_out))])
        self.irreps_out = Irreps(irreps_out)
        self.irreps_in = Irreps("1o") if irreps_in is None else Irreps(irreps_in)
        self.normalize = normalize
        self.normalization = normalization
        self._ls = [ir.l for mul, ir in self.irreps_out for _ in range(mul)]
        self._lmax = max(self._ls)

    def forward(self, v: torch.Tensor) -> torch.Tensor:
        if self.normalize:
            v = torch.nn.functional.normalize(v, dim=-1)
        sh = _raw_sh(self._lmax, v[..., 0], v[..., 1], v[..., 2])
        parts = []
        for l in self._ls:  # noqa: E741
            s = sh[l]
            if self.normalization == "component":
                s = s * math.sqrt(2 * l + 1)
            elif self.normalization == "integral":
                s = s * (math.sqrt(2 * l + 1) / math.sqrt(4 * math.pi))
            parts.append(s)
        return torch.cat(parts, dim=-1)
