"""Generalised Clebsch-Gordan tensors U_nu of the MACE product basis — CPU restatement.

TEST INFRASTRUCTURE ONLY (oracle): only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.

Restates ``U_matrix_real`` / ``_wigner_nj`` (reference ``src/physicsml/models/mace/modules/cg.py:19-133``) iteratively:
level n couples every level-(n-1) tensor with each input irrep through w3j(l_out, l_left, l)*sqrt(2 l_out+1)
(``cg.py:54-65``), embeds the result in the full S-wide last axis (``cg.py:66-73``) and keeps the level sorted
(stably) by output irrep (``cg.py:89``).  ``u_matrix`` then stacks all tensors that end in the requested irrep
on a trailing axis, squeezed the way ``cg.py:117-131`` + ``symmetric_contraction.py:97-105`` do.

Pinned against the ``U_matrices.{1,2}`` tensors of the reference fixture checkpoint (max |diff| = 0.0) and against the
reference's own ``cg.py`` imported unmodified (tests/test_oracle_pins.py).
"""
from __future__ import annotations

import functools
import os
import sys

import torch

_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")
if _SHIMS not in sys.path:
    sys.path.insert(0, _SHIMS)

from e3nn.o3 import Irrep, wigner_3j  # noqa: E402  (the oracle's own restatement, see oracle/shims/e3nn)


def _levels(irs: tuple, nu: int) -> list:
    """List of (irrep, tensor[2l+1, S, ..(n times).., S]) for all n-fold couplings, sorted by irrep."""
    dims = [ir.dim for ir in irs]
    S = sum(dims)
    offs = [sum(dims[:a]) for a in range(len(irs))]
    eye = torch.eye(S)
    level = [(ir, eye[offs[a] : offs[a] + ir.dim]) for a, ir in enumerate(irs)]
    for n in range(2, nu + 1):
        nxt = []
        for ir_left, c_left in level:
            for a, ir in enumerate(irs):
                for ir_out in ir_left * ir:
                    c = wigner_3j(ir_out.l, ir_left.l, ir.l) * ir_out.dim**0.5
                    c = torch.einsum("jk,ijl->ikl", c_left.flatten(1), c)
                    full = torch.zeros(ir_out.dim, *([S] * (n - 1)), S)
                    full[..., offs[a] : offs[a] + ir.dim] = c.reshape(ir_out.dim, *([S] * (n - 1)), ir.dim)
                    nxt.append((ir_out, full))
        level = sorted(nxt, key=lambda t: t[0])  # stable
    return level


@functools.lru_cache(maxsize=None)
def _u_matrix_cached(irs: tuple, out: tuple, nu: int) -> torch.Tensor:
    if nu > 3:
        raise NotImplementedError("correlation 4 applies an intermediate-irrep filter (cg.py:100-114); not restated")
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float32)  # the reference builds U in the default dtype (fp32)
    try:
        mats = [c.squeeze().unsqueeze(-1) for ir, c in _levels(tuple(Irrep(i) for i in irs), nu) if ir == Irrep(out)]
        u = torch.cat(mats, dim=-1)
    finally:
        torch.set_default_dtype(old)
    while u.dim() <= nu:  # symmetric_contraction.py:104-105
        u = u.unsqueeze(0)
    return u


def u_matrix(coupling_irreps, irrep_out, nu: int) -> torch.Tensor:
    """U_nu with shape [2L+1, S, ...(nu)..., S, k]; for L = 0 the leading axis is squeezed away: [S, ..., S, k]."""
    irs = tuple(tuple(Irrep(ir)) for ir in coupling_irreps)
    return _u_matrix_cached(irs, tuple(Irrep(irrep_out)), int(nu)).clone()
