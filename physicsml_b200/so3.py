"""Host-side (numpy, fp64) SO(3) tables the kernels are generated from: irreps bookkeeping, real Wigner-3j tensors
in e3nn's basis, channelwise tensor-product path tables and the generalised Clebsch-Gordan tensors of the MACE
product basis.

These are *plan-building* utilities: everything here runs once per model configuration on the host; the results
are baked into generated CUDA code (``physicsml_b200/csrc/gen``) or uploaded as small device tables.

Conventions follow e3nn==0.5.1 as used by the reference:
  * irreps (l, p) ordered 0o < 0e < 1o < ... ; a feature vector is a concatenation of blocks [mul, 2l+1], mul-major;
  * wigner_3j(l1,l2,l3) = Re(Q1 (x) Q2 (x) conj(Q3)^T . su2_cg), Frobenius-normalised;
  * 'uvu' tensor-product paths enumerated as in ``models/mace/modules/irreps_tools.py:14-43`` and
    ``models/nequip/modules/interaction_block.py:29-49`` of the reference;
  * generalised CG tensors U_nu as in ``models/mace/modules/cg.py:19-133``.
"""
from __future__ import annotations

import functools
import itertools
import re
from dataclasses import dataclass
from fractions import Fraction
from math import factorial

import numpy as np


# --------------------------------------------------------------------------------------------------------------------
# irreps
# --------------------------------------------------------------------------------------------------------------------
def parse_irreps(spec) -> list[tuple[int, int, int]]:
    """'128x0e + 128x1o' -> [(128, 0, +1), (128, 1, -1)].  Also accepts an iterable of (mul, (l, p)) / (mul, l, p)."""
    if isinstance(spec, str):
        out = []
        for part in spec.split("+"):
            part = part.strip()
            if not part:
                continue
            m = re.fullmatch(r"(?:(\d+)\s*x\s*)?(\d+)([eo])", part)
            if m is None:
                raise ValueError(f"cannot parse irreps term {part!r}")
            out.append((int(m.group(1) or 1), int(m.group(2)), 1 if m.group(3) == "e" else -1))
        return out
    out = []
    for item in spec:
        if len(item) == 3:
            out.append((int(item[0]), int(item[1]), int(item[2])))
        else:
            mul, ir = item
            out.append((int(mul), int(ir[0]), int(ir[1])))
    return out


def irreps_str(irreps) -> str:
    return "+".join(f"{m}x{l}{'e' if p == 1 else 'o'}" for m, l, p in irreps)


def irreps_dim(irreps) -> int:
    return sum(m * (2 * l + 1) for m, l, _ in irreps)


def irreps_offsets(irreps) -> list[int]:
    off, out = 0, []
    for m, l, _ in irreps:
        out.append(off)
        off += m * (2 * l + 1)
    return out


def sort_irreps(irreps):
    """Stable sort by (l, p) with p=-1 before p=+1; returns (sorted, perm) with perm[i_old] = i_new."""
    order = sorted(range(len(irreps)), key=lambda i: ((irreps[i][1], irreps[i][2]), i))
    perm = [0] * len(irreps)
    for new, old in enumerate(order):
        perm[old] = new
    return [irreps[i] for i in order], perm


def simplify_irreps(irreps):
    out: list[list[int]] = []
    for m, l, p in irreps:
        if out and out[-1][1] == l and out[-1][2] == p:
            out[-1][0] += m
        elif m > 0:
            out.append([m, l, p])
    return [tuple(x) for x in out]


def sh_irreps(lmax: int, parity: int = -1):
    return [(1, l, parity**l) for l in range(lmax + 1)]


# --------------------------------------------------------------------------------------------------------------------
# Wigner 3j (real basis)
# --------------------------------------------------------------------------------------------------------------------
def _su2_cg(j1: int, j2: int, j3: int) -> np.ndarray:
    f = factorial
    out = np.zeros((2 * j1 + 1, 2 * j2 + 1, 2 * j3 + 1))
    for m1 in range(-j1, j1 + 1):
        for m2 in range(-j2, j2 + 1):
            m3 = m1 + m2
            if abs(m3) > j3:
                continue
            pref = Fraction(
                (2 * j3 + 1) * f(j3 + j1 - j2) * f(j3 - j1 + j2) * f(j1 + j2 - j3) * f(j3 + m3) * f(j3 - m3),
                f(j1 + j2 + j3 + 1) * f(j1 - m1) * f(j1 + m1) * f(j2 - m2) * f(j2 + m2),
            )
            s = Fraction(0)
            for v in range(max(-j1 + j2 + m3, -j1 + m1, 0), min(j2 + j3 + m1, j3 - j1 + j2, j3 + m3) + 1):
                s += (-1) ** (v + j2 + m2) * Fraction(
                    f(j2 + j3 + m1 - v) * f(j1 - m1 + v),
                    f(v) * f(j3 - j1 + j2 - v) * f(j3 + m3 - v) * f(v + j1 - j2 - m3),
                )
            out[j1 + m1, j2 + m2, j3 + m3] = float(pref) ** 0.5 * float(s)
    return out


def _real_to_complex(l: int) -> np.ndarray:
    q = np.zeros((2 * l + 1, 2 * l + 1), dtype=np.complex128)
    s = 1.0 / np.sqrt(2.0)
    for m in range(1, l + 1):
        q[l - m, l + m] = s
        q[l - m, l - m] = -1j * s
        q[l + m, l + m] = (-1) ** m * s
        q[l + m, l - m] = 1j * (-1) ** m * s
    q[l, l] = 1.0
    return (-1j) ** l * q


@functools.lru_cache(maxsize=None)
def wigner_3j(l1: int, l2: int, l3: int) -> np.ndarray:
    """Real-basis 3j tensor [2l1+1, 2l2+1, 2l3+1], unit Frobenius norm (fp64; read-only)."""
    if not abs(l2 - l3) <= l1 <= l2 + l3:
        raise ValueError("triangle inequality violated")
    c = np.einsum(
        "ij,kl,mn,ikn->jlm",
        _real_to_complex(l1),
        _real_to_complex(l2),
        np.conj(_real_to_complex(l3).T),
        _su2_cg(l1, l2, l3).astype(np.complex128),
    )
    assert np.abs(c.imag).max() < 1e-9
    c = np.ascontiguousarray(c.real)
    c /= np.linalg.norm(c)
    c.setflags(write=False)
    return c


# --------------------------------------------------------------------------------------------------------------------
# channelwise ('uvu') tensor-product path table
# --------------------------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class TPPath:
    i_in: int  # block index in the node irreps
    l1: int
    l2: int  # index (= l) into the spherical-harmonic irreps
    l3: int
    p3: int
    i_out: int  # block index in the SORTED output irreps
    in_off: int  # per-channel offset of the in-block:  h[.., in_base + u*(2l1+1) + i]; in_off = sum_{b<i_in}(2l_b+1)
    out_off: int  # per-channel offset of the out-block (sorted order): sum over earlier sorted blocks of (2l+1)
    weight: float  # sqrt(2 l3 + 1)  (component / element normalisation)


@dataclass(frozen=True)
class TPSpec:
    """Everything the fused edge kernel needs to know about one channelwise tensor product, for uniform mul C."""

    node_ls: tuple  # ((l, p), ...) of the gathered node features, one block each with multiplicity C
    sh_lmax: int
    sh_parity: int  # -1: Y_l has parity (-1)^l, +1: all even
    paths: tuple  # TPPath in INSTRUCTION order (= weight-column order: column p*C + u)
    out_ls: tuple  # ((l, p), ...) sorted output blocks, one per path
    d_in: int  # per-channel input width  sum (2 l1 + 1)
    d_out: int  # per-channel output width sum_p (2 l3 + 1)
    n_sh: int  # (lmax+1)^2

    @property
    def n_paths(self) -> int:
        return len(self.paths)

    @property
    def key(self) -> str:
        """Signature used to pick the generated kernel."""
        n = "_".join(f"{l}{'e' if p == 1 else 'o'}" for l, p in self.node_ls)
        t = "_".join(f"{q.i_in}{q.l2}{q.l3}{'e' if q.p3 == 1 else 'o'}" for q in self.paths)
        return f"n{n}__sh{self.sh_lmax}{'o' if self.sh_parity == -1 else 'e'}__p{t}"


def build_tp_spec(node_irreps, sh_lmax: int, target_irreps, sh_parity: int = -1) -> TPSpec:
    node = parse_irreps(node_irreps)
    muls = {m for m, _, _ in node}
    if len(muls) != 1:
        raise ValueError(f"the fused edge kernel needs one multiplicity for all node blocks, got {irreps_str(node)}")
    target = {(l, p) for _, l, p in parse_irreps(target_irreps)}
    raw = []  # (i_in, l1, l2, l3, p3)
    for i, (_, l1, p1) in enumerate(node):
        for l2 in range(sh_lmax + 1):
            p2 = sh_parity**l2
            for l3 in range(abs(l1 - l2), l1 + l2 + 1):
                if (l3, p1 * p2) in target:
                    raw.append((i, l1, l2, l3, p1 * p2))
    blocks = [(1, l3, p3) for (_, _, _, l3, p3) in raw]
    sorted_blocks, perm = sort_irreps(blocks)
    out_off_sorted, acc = [], 0
    for _, l, _ in sorted_blocks:
        out_off_sorted.append(acc)
        acc += 2 * l + 1
    in_off, a = [], 0
    for _, l, _ in node:
        in_off.append(a)
        a += 2 * l + 1
    paths = tuple(
        TPPath(i, l1, l2, l3, p3, perm[k], in_off[i], out_off_sorted[perm[k]], float(np.sqrt(2 * l3 + 1)))
        for k, (i, l1, l2, l3, p3) in enumerate(raw)
    )
    return TPSpec(
        node_ls=tuple((l, p) for _, l, p in node),
        sh_lmax=sh_lmax,
        sh_parity=sh_parity,
        paths=paths,
        out_ls=tuple((l, p) for _, l, p in sorted_blocks),
        d_in=a,
        d_out=acc,
        n_sh=(sh_lmax + 1) ** 2,
    )


def tp_terms(spec: TPSpec):
    """Sparse list of (path_index, out_slot, in_slot, sh_slot, coef) with
    out[out_slot] += R[path] * coef * h[in_slot] * Y[sh_slot]   (per channel; coef includes sqrt(2 l3+1))."""
    terms = []
    for p, q in enumerate(spec.paths):
        w = wigner_3j(q.l1, q.l2, q.l3) * q.weight
        for i, j, k in itertools.product(range(2 * q.l1 + 1), range(2 * q.l2 + 1), range(2 * q.l3 + 1)):
            c = w[i, j, k]
            if abs(c) > 1e-12:
                terms.append((p, q.out_off + k, q.in_off + i, q.l2 * q.l2 + j, float(c)))
    return terms


# --------------------------------------------------------------------------------------------------------------------
# generalised Clebsch-Gordan tensors of the product basis
# --------------------------------------------------------------------------------------------------------------------
@functools.lru_cache(maxsize=None)
def u_matrix(coupling: tuple, out: tuple, nu: int) -> np.ndarray:
    """U_nu[M, i_1..i_nu, k] (fp64) for coupling irreps ((l,p),...) (one copy each) into irrep `out`.

    Always carries the leading M axis (size 2L+1), unlike the reference which squeezes it for L=0.
    """
    if nu > 3:
        raise NotImplementedError("correlation order > 3 is not supported")
    dims = [2 * l + 1 for l, _ in coupling]
    S = sum(dims)
    offs = np.concatenate([[0], np.cumsum(dims)])[:-1]
    eye = np.eye(S)
    level = [((l, p), eye[offs[a] : offs[a] + 2 * l + 1]) for a, (l, p) in enumerate(coupling)]
    for n in range(2, nu + 1):
        nxt = []
        for (ll, pl), cl in level:
            flat = cl.reshape(cl.shape[0], -1)
            for a, (l, p) in enumerate(coupling):
                for lo in range(abs(ll - l), ll + l + 1):
                    w = wigner_3j(lo, ll, l) * np.sqrt(2 * lo + 1)
                    c = np.einsum("jk,ijl->ikl", flat, w)
                    full = np.zeros((2 * lo + 1,) + (S,) * (n - 1) + (S,))
                    full[..., offs[a] : offs[a] + 2 * l + 1] = c.reshape((2 * lo + 1,) + (S,) * (n - 1) + (2 * l + 1,))
                    nxt.append(((lo, pl * p), full))
        order = sorted(range(len(nxt)), key=lambda i: (nxt[i][0], i))
        level = [nxt[i] for i in order]
    mats = [c for ir, c in level if ir == tuple(out)]
    if not mats:
        raise ValueError(f"irrep {out} cannot be reached at correlation {nu}")
    u = np.stack(mats, axis=-1)
    u.setflags(write=False)
    return u


def symmetrise(u: np.ndarray, nu: int) -> np.ndarray:
    """Average U[M, i_1..i_nu, k] over permutations of the nu feature axes (the contraction only sees this part)."""
    if nu == 1:
        return u
    acc = np.zeros_like(u)
    perms = list(itertools.permutations(range(nu)))
    for perm in perms:
        acc += np.transpose(u, (0,) + tuple(1 + a for a in perm) + (nu + 1,))
    return acc / len(perms)


def contraction_monomials(coupling: tuple, out: tuple, nu: int, tol: float = 1e-10):
    """{(k, M, (i_1<=..<=i_nu)): coef}: out_M = sum_k W_k sum_mono coef * x_{i_1}...x_{i_nu}  (one degree nu)."""
    u = symmetrise(u_matrix(coupling, out, nu), nu)
    res = {}
    K = u.shape[-1]
    for idx in itertools.combinations_with_replacement(range(u.shape[1]), nu):
        # multiplicity = number of distinct permutations of idx
        counts = {}
        for a in idx:
            counts[a] = counts.get(a, 0) + 1
        mult = factorial(nu)
        for cnt in counts.values():
            mult //= factorial(cnt)
        for M in range(u.shape[0]):
            for k in range(K):
                c = u[(M,) + idx + (k,)] * mult
                if abs(c) > tol:
                    res[(k, M, idx)] = float(c)
    return res


@functools.lru_cache(maxsize=None)
def reduced_contraction(coupling: tuple, out: tuple, nu: int, tol: float = 1e-9):
    """The K weight rows of one (output irrep, degree nu) of the symmetric contraction multiply K polynomials
    P_k[M](x) that are far from independent once the generalised-CG tensor is symmetrised over its nu feature axes
    (lmax 3, 1o, nu 3: 51 rows, rank 12).  Pick J = rank independent rows (sparsest first) and write every row in
    that basis, P_k = sum_j B[k, j] P_sel[j]; then

        out_M = sum_k W_k P_k[M](x) = sum_j W'_j P_sel[j][M](x),     W' = B^T W   (a tiny matrix product on the weights)

    so the kernels evaluate J instead of K polynomials (about a quarter of the monomial terms).
    Returns (monomials {(j, M, idx): coef} of the basis rows, B as a [K, J] float64 array, sel)."""
    monos = contraction_monomials(coupling, out, nu)
    K = u_matrix(coupling, out, nu).shape[-1]
    cols = sorted({(M, idx) for (_, M, idx) in monos})
    col_of = {c: i for i, c in enumerate(cols)}
    A = np.zeros((K, len(cols)))
    for (k, M, idx), c in monos.items():
        A[k, col_of[(M, idx)]] = c
    nnz = (np.abs(A) > 0).sum(axis=1)
    order = sorted(range(K), key=lambda k: (nnz[k], k))
    sel, basis = [], []  # chosen rows and their orthonormalised copies (modified Gram-Schmidt rank test)
    for k in order:
        if nnz[k] == 0:
            continue
        r = A[k].copy()
        for q in basis:
            r -= (r @ q) * q
        if np.linalg.norm(r) > tol * max(1.0, np.linalg.norm(A[k])):
            sel.append(k)
            basis.append(r / np.linalg.norm(r))
    sel.sort()  # a full-rank set reduces to the identity
    J = len(sel)
    if J == 0:
        return {}, np.zeros((K, 0)), ()
    B = np.linalg.lstsq(A[sel].T, A.T, rcond=None)[0].T  # A = B @ A[sel]
    if np.abs(B @ A[sel] - A).max() > 1e-10:
        raise AssertionError("basis reduction of the contraction polynomials is not exact")
    B[np.abs(B) < 1e-12] = 0.0
    for j, k in enumerate(sel):
        B[k] = 0.0
        B[k, j] = 1.0
    red = {(j, M, idx): c for j, k in enumerate(sel) for (kk, M, idx), c in monos.items() if kk == k}
    return red, B, tuple(sel)
