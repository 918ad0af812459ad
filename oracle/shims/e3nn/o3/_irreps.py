"""Irreps algebra restated from the published behaviour of e3nn==0.5.1 (``e3nn.o3.Irrep`` / ``e3nn.o3.Irreps``).

TEST INFRASTRUCTURE ONLY (oracle).  e3nn itself is not vendored under /root/reference
(pinned in ``pinned-versions/3.11/lockfile.core.txt:115``); the semantics restated here are the
ones listed in SURVEY.md Appendix A.1 and exercised by the reference at
``src/physicsml/models/mace/modules/mace.py:42-73``, ``irreps_tools.py:14-43`` and ``cg.py:19-133``.
"""
from __future__ import annotations

import collections
import re
from typing import Iterator


class Irrep(tuple):
    """(l, p) with p=+1 even / p=-1 odd.  Orders as a plain tuple: 0o < 0e < 1o < 1e ..."""

    def __new__(cls, l, p=None):  # noqa: E741
        if p is None:
            if isinstance(l, Irrep):
                return l
            if isinstance(l, _MulIr):
                return l.ir
            if isinstance(l, str):
                m = re.fullmatch(r"\s*(\d+)([eoy])\s*", l)
                if m is None:
                    raise ValueError(f"cannot parse irrep {l!r}")
                ll = int(m.group(1))
                p = {"e": 1, "o": -1, "y": (-1) ** ll}[m.group(2)]
                l = ll  # noqa: E741
            elif isinstance(l, tuple):
                l, p = l  # noqa: E741
        if not isinstance(l, int) or l < 0:
            raise ValueError(f"l must be a non-negative integer, got {l!r}")
        if p not in (-1, 1):
            raise ValueError(f"parity must be +1 or -1, got {p!r}")
        return super().__new__(cls, (l, p))

    @property
    def l(self) -> int:  # noqa: E743
        return self[0]

    @property
    def p(self) -> int:
        return self[1]

    @property
    def dim(self) -> int:
        return 2 * self[0] + 1

    def __repr__(self) -> str:
        return f"{self[0]}{'e' if self[1] == 1 else 'o'}"

    def __mul__(self, other) -> Iterator["Irrep"]:
        other = Irrep(other)
        p = self.p * other.p
        for l in range(abs(self.l - other.l), self.l + other.l + 1):  # noqa: E741
            yield Irrep(l, p)

    def __rmul__(self, mul: int) -> "Irreps":
        return Irreps([(mul, self)])

    def __add__(self, other):
        return Irreps(self) + Irreps(other)

    def is_scalar(self) -> bool:
        return self.l == 0 and self.p == 1


class _MulIr(tuple):
    def __new__(cls, mul, ir=None):
        if ir is None:
            mul, ir = mul
        return super().__new__(cls, (int(mul), Irrep(ir)))

    @property
    def mul(self) -> int:
        return self[0]

    @property
    def ir(self) -> Irrep:
        return self[1]

    @property
    def dim(self) -> int:
        return self[0] * self[1].dim

    def __repr__(self) -> str:
        return f"{self[0]}x{self[1]}"


_SortRet = collections.namedtuple("sort", ["irreps", "p", "inv"])


class Irreps(tuple):
    def __new__(cls, irreps=None):
        if isinstance(irreps, Irreps):
            return super().__new__(cls, irreps)
        out = []
        if irreps is None:
            pass
        elif isinstance(irreps, Irrep):
            out.append(_MulIr(1, irreps))
        elif isinstance(irreps, _MulIr):
            out.append(irreps)
        elif isinstance(irreps, str):
            if irreps.strip() != "":
                for part in irreps.split("+"):
                    part = part.strip()
                    if "x" in part:
                        mul, ir = part.split("x")
                        out.append(_MulIr(int(mul), Irrep(ir)))
                    else:
                        out.append(_MulIr(1, Irrep(part)))
        else:
            for item in irreps:
                if isinstance(item, _MulIr):
                    out.append(item)
                elif isinstance(item, Irrep):
                    out.append(_MulIr(1, item))
                elif isinstance(item, str):
                    out.extend(Irreps(item))
                elif len(item) == 2 and not isinstance(item[0], (tuple, str)) and isinstance(item[1], (Irrep, str, tuple)):
                    out.append(_MulIr(item[0], Irrep(item[1])))
                else:
                    out.append(_MulIr(1, Irrep(item)))
        return super().__new__(cls, out)

    @staticmethod
    def spherical_harmonics(lmax: int, p: int = -1) -> "Irreps":
        return Irreps([(1, (l, p**l)) for l in range(lmax + 1)])

    @property
    def dim(self) -> int:
        return sum(mi.dim for mi in self)

    @property
    def num_irreps(self) -> int:
        return sum(mi.mul for mi in self)

    @property
    def ls(self) -> list[int]:
        return [mi.ir.l for mi in self for _ in range(mi.mul)]

    @property
    def lmax(self) -> int:
        if len(self) == 0:
            raise ValueError("Cannot get lmax of empty Irreps")
        return max(mi.ir.l for mi in self)

    def slices(self) -> list[slice]:
        s, i = [], 0
        for mi in self:
            s.append(slice(i, i + mi.dim))
            i += mi.dim
        return s

    def count(self, ir) -> int:
        ir = Irrep(ir)
        return sum(mi.mul for mi in self if mi.ir == ir)

    def __contains__(self, ir) -> bool:
        ir = Irrep(ir)
        return any(mi.ir == ir for mi in self)

    def __getitem__(self, i):
        x = super().__getitem__(i)
        if isinstance(i, slice):
            return Irreps(x)
        return x

    def __add__(self, other) -> "Irreps":
        return Irreps(tuple.__add__(self, Irreps(other)))

    def __mul__(self, n: int) -> "Irreps":
        return Irreps(tuple.__mul__(self, n))

    __rmul__ = __mul__

    def simplify(self) -> "Irreps":
        out: list[tuple[int, Irrep]] = []
        for mul, ir in self:
            if out and out[-1][1] == ir:
                out[-1] = (out[-1][0] + mul, ir)
            elif mul > 0:
                out.append((mul, ir))
        return Irreps(out)

    def remove_zero_multiplicities(self) -> "Irreps":
        return Irreps([(mul, ir) for mul, ir in self if mul > 0])

    def sort(self) -> _SortRet:
        # stable sort by (ir, original index); p[i_old] = i_new, inv[i_new] = i_old
        order = sorted((ir, i, mul) for i, (mul, ir) in enumerate(self))
        inv = tuple(i for _, i, _ in order)
        p = [0] * len(inv)
        for new, old in enumerate(inv):
            p[old] = new
        return _SortRet(Irreps([(mul, ir) for ir, _, mul in order]), tuple(p), inv)

    def __repr__(self) -> str:
        return "+".join(repr(mi) for mi in self)
