"""Real-basis Wigner 3j symbols, restated from the published algorithm of e3nn==0.5.1
(``e3nn.o3.wigner_3j``).  TEST INFRASTRUCTURE ONLY (oracle).

Pinned bit-exactly against the ten ``_w3j_*`` buffers stored in the reference's fixture
checkpoints (``tests/data/{mace,nequip}_model/model_artefacts/module_checkpoint.ckpt``);
see ``tests/test_oracle_pins.py``.
"""
from __future__ import annotations

import functools
from fractions import Fraction
from math import factorial

import torch


def _su2_cg_coeff(j1: int, m1: int, j2: int, m2: int, j3: int, m3: int) -> float:
    """Condon-Shortley <j1 m1 j2 m2 | j3 m3> (Racah's closed form, exact rational under the root)."""
    if m3 != m1 + m2:
        return 0.0
    vmin = max(-j1 + j2 + m3, -j1 + m1, 0)
    vmax = min(j2 + j3 + m1, j3 - j1 + j2, j3 + m3)

    def f(n: int) -> int:
        return factorial(round(n))

    c = (
        Fraction(
            (2 * j3 + 1) * f(j3 + j1 - j2) * f(j3 - j1 + j2) * f(j1 + j2 - j3) * f(j3 + m3) * f(j3 - m3),
            f(j1 + j2 + j3 + 1) * f(j1 - m1) * f(j1 + m1) * f(j2 - m2) * f(j2 + m2),
        )
    ) ** 0.5
    s = Fraction(0)
    for v in range(vmin, vmax + 1):
        s += (-1) ** (v + j2 + m2) * Fraction(
            f(j2 + j3 + m1 - v) * f(j1 - m1 + v),
            f(v) * f(j3 - j1 + j2 - v) * f(j3 + m3 - v) * f(v + j1 - j2 - m3),
        )
    return float(c * s)


def _su2_cg(j1: int, j2: int, j3: int) -> torch.Tensor:
    mat = torch.zeros(2 * j1 + 1, 2 * j2 + 1, 2 * j3 + 1, dtype=torch.float64)
    if abs(j1 - j2) <= j3 <= j1 + j2:
        for m1 in range(-j1, j1 + 1):
            for m2 in range(-j2, j2 + 1):
                if abs(m1 + m2) <= j3:
                    mat[j1 + m1, j2 + m2, j3 + m1 + m2] = _su2_cg_coeff(j1, m1, j2, m2, j3, m1 + m2)
    return mat


def _real_to_complex(l: int) -> torch.Tensor:  # noqa: E741
    q = torch.zeros(2 * l + 1, 2 * l + 1, dtype=torch.complex128)
    s = 2**-0.5
    for m in range(-l, 0):
        q[l + m, l + abs(m)] = s
        q[l + m, l - abs(m)] = -1j * s
    q[l, l] = 1
    for m in range(1, l + 1):
        q[l + m, l + abs(m)] = (-1) ** m * s
        q[l + m, l - abs(m)] = 1j * (-1) ** m * s
    return (-1j) ** l * q


@functools.lru_cache(maxsize=None)
def _so3_cg(l1: int, l2: int, l3: int) -> torch.Tensor:
    q1, q2, q3 = _real_to_complex(l1), _real_to_complex(l2), _real_to_complex(l3)
    c = _su2_cg(l1, l2, l3).to(torch.complex128)
    c = torch.einsum("ij,kl,mn,ikn->jlm", q1, q2, torch.conj(q3.T), c)
    assert c.imag.abs().max() < 1e-5
    c = c.real
    return c / c.norm()


def wigner_3j(l1: int, l2: int, l3: int, dtype=None, device=None) -> torch.Tensor:
    assert abs(l2 - l3) <= l1 <= l2 + l3
    if dtype is None:
        dtype = torch.get_default_dtype()
    return _so3_cg(int(l1), int(l2), int(l3)).to(dtype=dtype, device=device).clone()
