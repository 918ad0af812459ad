"""Generates the straight-line CUDA device code for the two kernels whose arithmetic is a sparse, statically known
multilinear form:

  * the channelwise edge tensor product (reference ``models/mace/modules/blocks.py:164-171,214-218``,
    ``models/nequip/modules/interaction_block.py:51-58,95``): per (edge, channel)
        out[o] += R[p] * c * h[i] * Y[j]                                     (sparse Wigner-3j, c includes sqrt(2 l3+1))
  * the symmetric contraction (row a12; reference ``models/mace/modules/symmetric_contraction.py:218-239``): per (node, channel)
        out[M] = sum_nu sum_k W_nu[z,k,c] * sum_mono c * x_i1 .. x_inu        (symmetrised generalised CG)

Register indices must be compile-time constants, so the path / monomial tables are unrolled into source here rather
than walked at run time.  Output: one header per kernel family under ``physicsml_b200/csrc/gen`` containing
explicit template specialisations ``TPCode<ID>`` / ``SymCode<ID>``, plus ``registry.json`` mapping signature -> ID.
"""
from __future__ import annotations

import json
import os
from collections import defaultdict

from . import so3

GEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "gen")


def _f(x: float) -> str:
    s = repr(float(x))
    if "e" in s or "." in s or "inf" in s or "nan" in s:
        return s + "f"
    return s + ".0f"


# ----------------------------------------------------------------------------------------------------------------------
# tensor product
# ----------------------------------------------------------------------------------------------------------------------
def default_tp_specs() -> list[so3.TPSpec]:
    """Signatures compiled by default: every BASELINE config + the reference's fixture models (+ small test shapes)."""
    specs = {}

    def add(node, lmax, target, parity=-1):
        s = so3.build_tp_spec(node, lmax, target, parity)
        specs[s.key] = s

    for lmax in (1, 2, 3):
        tgt = so3.irreps_str([(1, l, (-1) ** l) for l in range(lmax + 1)])
        add("1x0e", lmax, tgt)
        add("1x0e+1x1o", lmax, tgt)
        if lmax >= 2:
            add("1x0e+1x1o+1x2e", lmax, tgt)
    # NequIP: hidden = F x {l e} (+ {l o} when parity=True), restricted to what is reachable layer by layer
    # (nequip.py:68-100, convnet_layer.py:44-81); sh parity -1 (parity=True) or +1 (parity=False)
    for sh_parity in (-1, 1):
        for lmax in (1, 2):
            hidden = [(1, l, 1) for l in range(lmax + 1)]
            if sh_parity == -1:
                hidden += [(1, l, -1) for l in range(lmax + 1)]
            hidden, _ = so3.sort_irreps(hidden)
            node = [(1, 0, 1)]
            for _layer in range(4):
                reach = set()
                for _, l1, p1 in node:
                    for l2 in range(lmax + 1):
                        for l3 in range(abs(l1 - l2), l1 + l2 + 1):
                            reach.add((l3, p1 * sh_parity**l2))
                out = [(1, l, p) for _, l, p in hidden if (l, p) in reach]
                gate_ir = (0, 1) if (0, 1) in reach else (0, -1)
                tgt = sorted(set([(l, p) for _, l, p in out] + [gate_ir]))
                add(so3.irreps_str(node), lmax, so3.irreps_str([(1, l, p) for l, p in tgt]), sh_parity)
                node = out
    return list(specs.values())


def _emit_tp(spec: so3.TPSpec, ident: int) -> str:
    terms = so3.tp_terms(spec)
    NP, DIN, DOUT, NSH = spec.n_paths, spec.d_in, spec.d_out, spec.n_sh
    by_path = defaultdict(list)
    for p, o, i, j, c in terms:
        by_path[p].append((o, i, j, c))
    L = []
    L.append(f"// {spec.key}")
    L.append(f"template <> struct TPCode<{ident}> {{")
    L.append(f"  static constexpr int NP = {NP}, DIN = {DIN}, DOUT = {DOUT}, NSH = {NSH};")
    L.append(f"  // the whole signature seen as one path group (see TPGroup below)")
    L.append(f"  static constexpr int P0 = 0, NP_FULL = {NP}, H_OFF = 0, DIN_FULL = {DIN}, DOUT_FULL = {DOUT}, NGROUPS = 1;")
    # ---- per-channel gather / scatter of the irreps layout (blocks [C, 2l+1], mul-major)
    L.append("  // h_row points at node_feats[s, 0]; block b of irrep l sits at C*in_off_b + u*(2l+1) + i")
    L.append("  __device__ __forceinline__ static void load_h(const float* __restrict__ hrow, int C, int u, float (&h)[DIN]) {")
    off = 0
    for l, _ in spec.node_ls:
        d = 2 * l + 1
        for i in range(d):
            L.append(f"    h[{off + i}] = __ldg(hrow + (size_t)C * {off} + (size_t)u * {d} + {i});")
        off += d
    L.append("  }")
    L.append("  // same gather from a row staged in shared memory (pipelined kernels)")
    L.append("  __device__ __forceinline__ static void load_h_smem(const float* hrow, int C, int u, float (&h)[DIN]) {")
    off = 0
    for l, _ in spec.node_ls:
        d = 2 * l + 1
        for i in range(d):
            L.append(f"    h[{off + i}] = hrow[C * {off} + u * {d} + {i}];")
        off += d
    L.append("  }")
    L.append("  __device__ __forceinline__ static void atomic_add_h(float* __restrict__ hrow, int C, int u, const float (&h)[DIN]) {")
    off = 0
    for l, _ in spec.node_ls:
        d = 2 * l + 1
        for i in range(d):
            L.append(f"    atomicAdd(hrow + (size_t)C * {off} + (size_t)u * {d} + {i}, h[{off + i}]);")
        off += d
    L.append("  }")
    # output row: block b of irrep l at C*off_b; inside a block either the e3nn order u*(2l+1) + m (MM = false) or
    # component-major m*C + u (MM = true: unit stride along the channels, which is what both the packed stores of this
    # kernel and the k-loop of the GEMM that consumes the row want)
    for name, op in (("load_out", "a[{k}] = __ldg(row + (size_t)C * {off} + (MM ? (size_t){m} * C + u : (size_t)u * {d} + {m}));"),
                     ("store_out", "row[(size_t)C * {off} + (MM ? (size_t){m} * C + u : (size_t)u * {d} + {m})] = a[{k}];")):
        const = "const " if name == "load_out" else ""
        aconst = "" if name == "load_out" else "const "
        L.append(f"  template <bool MM> __device__ __forceinline__ static void {name}({const}float* __restrict__ row, int C, int u, {aconst}float (&a)[DOUT]) {{")
        off = 0
        for l, _ in spec.out_ls:
            d = 2 * l + 1
            for m in range(d):
                L.append("    " + op.format(k=off + m, off=off, d=d, m=m))
            off += d
        L.append("  }")
    # ---- forward: acc[o] += R[p] * sum c h[i] Y[j]
    L.append("  __device__ __forceinline__ static void fwd(const float (&h)[DIN], const float (&Y)[NSH], const float (&R)[NP], float (&acc)[DOUT]) {")
    for p in range(NP):
        grp = defaultdict(list)
        for o, i, j, c in by_path[p]:
            grp[o].append((i, j, c))
        L.append("    {")
        for o in sorted(grp):
            expr = " + ".join(f"{_f(c)} * (h[{i}] * Y[{j}])" for i, j, c in grp[o])
            L.append(f"      acc[{o}] = fmaf(R[{p}], {expr}, acc[{o}]);")
        L.append("    }")
    L.append("  }")
    # ---- backward: given g = dL/d out (receiver row), accumulate dh, dY; write dR
    L.append("  template <bool WANT_H, bool WANT_Y, bool WANT_R, int NY>")
    L.append("  __device__ __forceinline__ static void bwd(const float (&h)[DIN], const float (&Y)[NSH], const float (&R)[NP], const float (&g)[DOUT],")
    L.append("                                             float (&dh)[DIN], float (&dY)[NY], float (&dR)[NP]) {")
    for p in range(NP):
        outs = sorted({o for o, _, _, _ in by_path[p]})
        L.append("    {")
        L.append("      if (WANT_R) {")
        expr = " + ".join(f"{_f(c)} * (h[{i}] * Y[{j}]) * g[{o}]" for o, i, j, c in by_path[p])
        L.append(f"        dR[{p}] = {expr};")
        L.append("      }")
        L.append("      if (WANT_H || WANT_Y) {")
        for o in outs:
            L.append(f"        const float rg{o} = R[{p}] * g[{o}];")
        L.append("        if (WANT_H) {")
        gi = defaultdict(list)
        for o, i, j, c in by_path[p]:
            gi[i].append((o, j, c))
        for i in sorted(gi):
            expr = " + ".join(f"{_f(c)} * (Y[{j}] * rg{o})" for o, j, c in gi[i])
            L.append(f"          dh[{i}] += {expr};")
        L.append("        }")
        L.append("        if (WANT_Y) {")
        gj = defaultdict(list)
        for o, i, j, c in by_path[p]:
            gj[j].append((o, i, c))
        for j in sorted(gj):
            expr = " + ".join(f"{_f(c)} * (h[{i}] * rg{o})" for o, i, c in gj[j])
            L.append(f"          dY[{j}] += {expr};")
        L.append("        }")
        L.append("      }")
        L.append("    }")
    L.append("  }")
    _emit_tp_packed(spec, by_path, L, list(range(NP)), 1)
    L.append("};")
    groups = tp_path_groups(spec)
    L.append(f"template <> struct TPGroups<{ident}> {{ static constexpr int N = {len(groups)}; }};")
    if len(groups) == 1:
        L.append(f"template <> struct TPGroup<{ident}, 0> : TPCode<{ident}> {{}};")
    else:
        for g, paths in enumerate(groups):
            L.append(f"// {spec.key}: path group {g} = paths {paths[0]}..{paths[-1]}")
            L.append(f"template <> struct TPGroup<{ident}, {g}> {{")
            _emit_tp_packed(spec, by_path, L, paths, len(groups))
            L.append("};")
    return "\n".join(L)


# two channels of accumulators (or of the cotangent row) per thread must fit the register file: 48 is the hard limit of the
# pipelined kernels; a smaller value cuts narrower signatures into groups too (more warps per edge, fewer registers each)
PIPE_MAX_DOUT = int(os.environ.get("PML_TP_GROUP_DOUT", "48"))


def tp_path_groups(spec: so3.TPSpec) -> list:
    """contiguous path ranges whose output width is <= PIPE_MAX_DOUT each (balanced), so that signatures with wide
    output rows (MACE large L=2: 71 components; NequIP layers >= 2: 93 / 102) run the pipelined kernels group by group:
    every group reads its own contiguous slice of the R row and of the gathered h row and owns its output blocks"""
    widths = [2 * q.l3 + 1 for q in spec.paths]
    total = sum(widths)
    n_groups = max(1, -(-total // PIPE_MAX_DOUT))
    while True:
        target = total / n_groups
        groups, cur, acc = [], [], 0
        for p, w in enumerate(widths):
            if cur and len(groups) < n_groups - 1 and acc + w / 2 > target:
                groups.append(cur)
                cur, acc = [], 0
            cur.append(p)
            acc += w
        groups.append(cur)
        if all(sum(widths[p] for p in g) <= PIPE_MAX_DOUT for g in groups):
            return groups
        n_groups += 1


def _f32_is(c: float, v: float) -> bool:
    import numpy as np

    return float(np.float32(c)) == v


def _emit_tp_packed(spec: so3.TPSpec, by_path, L: list, paths: list, n_groups: int) -> None:
    """Two channels per thread on the packed fp32x2 pipe (FFMA2 / FMUL2, sm_100): every operand is a float2 holding
    the same quantity for channels (2t, 2t+1); Y arrives as duplicated pairs (Y_j, Y_j).  The arithmetic is factored
    to ~1 packed instruction per non-zero 3j entry and cotangent:
      fwd : rh_i = R_p h_i ;  acc_k += (|c| rh_i) (+-Y_j)
      bwd : a = |c| g_k ;  dh'_i += a Y_j ;  dY'_j += a h_i ;  then per path  dR_p = sum_i h_i dh'_i,
            dh_i += R_p dh'_i,  dY_j += R_p . dY'_j  (the two channels are summed here: dY is per edge).

    `paths` selects a contiguous PATH GROUP of the signature (all paths for the main TPCode struct).  A group sees
    compact local arrays: R / dR [len(paths)], the slice [H_OFF, H_OFF + DIN) of the gathered h row, and acc / g over
    the output blocks of its own paths only; row addressing (load_out2 / store_out2 / atomic_add_h2) uses the offsets
    of the FULL rows."""
    NP_FULL, DIN_FULL, DOUT_FULL, NSH = spec.n_paths, spec.d_in, spec.d_out, spec.n_sh
    grouped = n_groups > 1
    P0 = paths[0]
    ins_used = sorted({i for p in paths for _, i, _, _ in by_path[p]})
    H_OFF, H_END = 0, DIN_FULL
    if grouped:  # the slice of the h row from the first to the last node block this group reads (whole blocks)
        touched, off = [], 0
        for l, _ in spec.node_ls:
            d = 2 * l + 1
            if any(off <= i < off + d for i in ins_used):
                touched.append((off, off + d))
            off += d
        H_OFF, H_END = touched[0][0], touched[-1][1]
    DIN = H_END - H_OFF
    out_slots = sorted({o for p in paths for o, _, _, _ in by_path[p]}) if grouped else list(range(DOUT_FULL))
    omap = {o: k for k, o in enumerate(out_slots)}
    DOUT, NP = len(out_slots), len(paths)
    if grouped:
        L.append(f"  static constexpr int NP = {NP}, DIN = {DIN}, DOUT = {DOUT}, NSH = {NSH};")
        L.append(f"  static constexpr int P0 = {P0}, NP_FULL = {NP_FULL}, H_OFF = {H_OFF}, DIN_FULL = {DIN_FULL}, DOUT_FULL = {DOUT_FULL}, NGROUPS = {n_groups};")

    def scaled(cache, tag, c, base, lines, indent):
        """name of |c|*base (a float2 expression name), emitting the multiply once per distinct (|c|, base)"""
        a = abs(c)
        if _f32_is(a, 1.0):
            return base
        key = (float(__import__("numpy").float32(a)), base)
        if key not in cache:
            name = f"{tag}{len(cache)}"
            lines.append(f"{indent}const float2 {name} = f2mul(f2dup({_f(a)}), {base});")
            cache[key] = name
        return cache[key]

    # ---------------- forward
    L.append("  // packed: h, R, acc hold channels (2t, 2t+1); Y[j] = (Y_j, Y_j)")
    L.append("  __device__ __forceinline__ static void fwd2(const float2 (&h)[DIN], const float2 (&Y)[NSH], const float2 (&R)[NP], float2 (&acc)[DOUT]) {")
    for p in paths:
        terms = by_path[p]
        L.append("    {")
        ins = sorted({i for _, i, _, _ in terms})
        for i in ins:
            L.append(f"      const float2 rh{i} = f2mul(R[{p - P0}], h[{i - H_OFF}]);")
        cache = {}
        for o, i, j, c in terms:
            sname = scaled(cache, "s", c, f"rh{i}", L, "      ")
            y = f"Y[{j}]" if c > 0 else f"f2neg(Y[{j}])"
            L.append(f"      acc[{omap[o]}] = f2fma({sname}, {y}, acc[{omap[o]}]);")
        L.append("    }")
    L.append("  }")

    # ---------------- backward
    L.append("  // dY (per edge) receives the sum over this thread's two channels; dh, dR stay packed")
    L.append("  template <bool WANT_H, bool WANT_Y, bool WANT_R, int NY>")
    L.append("  __device__ __forceinline__ static void bwd2(const float2 (&h)[DIN], const float2 (&Y)[NSH], const float2 (&R)[NP], const float2 (&g)[DOUT],")
    L.append("                                              float2 (&dh)[DIN], float (&dY)[NY], float2 (&dR)[NP]) {")
    for p in paths:
        terms = by_path[p]
        ins = sorted({i for _, i, _, _ in terms})
        js = sorted({j for _, _, j, _ in terms})
        L.append("    {")
        cache = {}
        pre = []
        names = {}
        for o, i, j, c in terms:
            names[(o, i, j)] = scaled(cache, "a", c, f"g[{omap[o]}]", pre, "      ")
        L.extend(pre)
        L.append("      if (WANT_H || WANT_R) {")
        seen = set()
        for o, i, j, c in terms:
            y = f"Y[{j}]" if c > 0 else f"f2neg(Y[{j}])"
            a = names[(o, i, j)]
            if i not in seen:
                L.append(f"        float2 dhp{i} = f2mul({a}, {y});")
                seen.add(i)
            else:
                L.append(f"        dhp{i} = f2fma({a}, {y}, dhp{i});")
        L.append("        if (WANT_R) {")
        for n, i in enumerate(ins):
            if n == 0:
                L.append(f"          float2 r = f2mul(h[{i - H_OFF}], dhp{i});")
            else:
                L.append(f"          r = f2fma(h[{i - H_OFF}], dhp{i}, r);")
        L.append(f"          dR[{p - P0}] = r;")
        L.append("        }")
        L.append("        if (WANT_H) {")
        for i in ins:
            L.append(f"          dh[{i - H_OFF}] = f2fma(R[{p - P0}], dhp{i}, dh[{i - H_OFF}]);")
        L.append("        }")
        L.append("      }")
        L.append("      if (WANT_Y) {")
        seen = set()
        for o, i, j, c in terms:
            hh = f"h[{i - H_OFF}]" if c > 0 else f"f2neg(h[{i - H_OFF}])"
            a = names[(o, i, j)]
            if j not in seen:
                L.append(f"        float2 dyp{j} = f2mul({a}, {hh});")
                seen.add(j)
            else:
                L.append(f"        dyp{j} = f2fma({a}, {hh}, dyp{j});")
        for j in js:
            L.append(f"        dY[{j}] = fmaf(R[{p - P0}].x, dyp{j}.x, fmaf(R[{p - P0}].y, dyp{j}.y, dY[{j}]));")
        L.append("      }")
        L.append("    }")
    L.append("  }")

    # ---------------- packed row access: channels (u, u+1), u even.  hrow points at the group's slice of the staged row
    # (slot H_OFF of the full row sits at hrow[0]); the node blocks inside the slice keep their [C, 2l+1] layout
    L.append("  __device__ __forceinline__ static void load_h2_smem(const float* hrow, int C, int u, float2 (&h)[DIN]) {")
    off = 0
    for l, _ in spec.node_ls:
        d = 2 * l + 1
        for i in range(d):
            s_ = off + i
            if H_OFF <= s_ < H_END:
                L.append(f"    h[{s_ - H_OFF}] = make_float2(hrow[C * {off - H_OFF} + u * {d} + {i}], hrow[C * {off - H_OFF} + (u + 1) * {d} + {i}]);")
        off += d
    L.append("  }")
    # output blocks of the group's paths, addressed in the FULL row
    blocks = []
    off = 0
    for l, _ in spec.out_ls:
        d = 2 * l + 1
        if any((off + m) in omap for m in range(d)):
            blocks.append((off, d))
        off += d
    L.append("  template <bool MM> __device__ __forceinline__ static void load_out2(const float* __restrict__ row, int C, int u, float2 (&a)[DOUT]) {")
    for off, d in blocks:
        for m in range(d):
            k = omap[off + m]
            L.append(f"    if (MM) a[{k}] = __ldg(reinterpret_cast<const float2*>(row + (size_t)C * {off} + (size_t){m} * C + u));")
            L.append(f"    else a[{k}] = make_float2(__ldg(row + (size_t)C * {off} + (size_t)u * {d} + {m}), __ldg(row + (size_t)C * {off} + (size_t)(u + 1) * {d} + {m}));")
    L.append("  }")
    L.append("  template <bool MM> __device__ __forceinline__ static void store_out2(float* __restrict__ row, int C, int u, const float2 (&a)[DOUT]) {")
    for off, d in blocks:
        for m in range(d):
            k = omap[off + m]
            L.append(f"    if (MM) *reinterpret_cast<float2*>(row + (size_t)C * {off} + (size_t){m} * C + u) = a[{k}];")
            L.append("    else {")
            L.append(f"      row[(size_t)C * {off} + (size_t)u * {d} + {m}] = a[{k}].x;")
            L.append(f"      row[(size_t)C * {off} + (size_t)(u + 1) * {d} + {m}] = a[{k}].y;")
            L.append("    }")
    L.append("  }")
    # fp32 reductions into dh[s] (FULL row pointer): the 2(2l+1) floats of channels (u, u+1) of one block are contiguous
    # and 8-byte aligned (u even) -> (2l+1) vector reds instead of 2(2l+1) scalar ones; only the blocks this group feeds
    L.append("  __device__ __forceinline__ static void atomic_add_h2(float* __restrict__ hrow, int C, int u, const float2 (&h)[DIN], uint64_t pol) {")
    off = 0
    for l, _ in spec.node_ls:
        d = 2 * l + 1
        if all(H_OFF <= off + i < H_END for i in range(d)) and any((off + i) in ins_used for i in range(d)):
            flat = [f"h[{off + i - H_OFF}].x" for i in range(d)] + [f"h[{off + i - H_OFF}].y" for i in range(d)]
            for v in range(d):
                L.append(f"    red_add_f32x2(hrow + (size_t)C * {off} + (size_t)u * {d} + {2 * v}, {flat[2 * v]}, {flat[2 * v + 1]}, pol);")
        off += d
    L.append("  }")


# ----------------------------------------------------------------------------------------------------------------------
# symmetric contraction
# ----------------------------------------------------------------------------------------------------------------------
class SymSpec:
    """One ``SymmetricContraction``: coupling irreps 0..lmax (parity (-1)^l), output irreps, correlation order."""

    def __init__(self, lmax: int, outs, corr: int) -> None:
        self.lmax = lmax
        self.coupling = tuple((l, (-1) ** l) for l in range(lmax + 1))
        self.outs = tuple((int(l), int(p)) for l, p in outs)
        self.corr = int(corr)
        self.S = (lmax + 1) ** 2
        # weight rows per (out, nu): K_full as in the reference's parameters, K after the basis reduction of the
        # symmetrised polynomials (so3.reduced_contraction) - the kernels see W' = B^T W with K rows
        self.K_full = [[so3.u_matrix(self.coupling, o, nu).shape[-1] for nu in range(1, corr + 1)] for o in self.outs]
        self.K = [[so3.reduced_contraction(self.coupling, o, nu)[1].shape[1] for nu in range(1, corr + 1)] for o in self.outs]
        self.d_out = sum(2 * l + 1 for l, _ in self.outs)

    @property
    def key(self) -> str:
        o = "_".join(f"{l}{'e' if p == 1 else 'o'}" for l, p in self.outs)
        return f"l{self.lmax}__o{o}__c{self.corr}"


def default_sym_specs() -> list[SymSpec]:
    specs = {}
    for lmax in (2, 3):
        for corr in (2, 3):
            for outs in ([(0, 1)], [(0, 1), (1, -1)], [(0, 1), (1, -1), (2, 1)]):
                if max(l for l, _ in outs) > lmax:
                    continue
                s = SymSpec(lmax, outs, corr)
                specs[s.key] = s
    return list(specs.values())


def _mono_expr(idx, xs="x") -> str:
    return " * ".join(f"{xs}[{a}]" for a in idx)


def _dmono(idx, j):
    """d/dx_j of prod x_idx -> (multiplicity, remaining index tuple) or None."""
    cnt = idx.count(j)
    if cnt == 0:
        return None
    rest = list(idx)
    rest.remove(j)
    return cnt, tuple(rest)


def _emit_sym(spec: SymSpec, ident: int) -> str:
    S, corr = spec.S, spec.corr
    NO = len(spec.outs)
    L = []
    L.append(f"// {spec.key}")
    L.append(f"template <> struct SymCode<{ident}> {{")
    L.append(f"  static constexpr int S = {S}, DOUT = {spec.d_out}, NOUT = {NO}, CORR = {corr};")
    ks = ", ".join("{" + ", ".join(str(k) for k in row) + "}" for row in spec.K)
    L.append(f"  // K[out][nu-1]: {ks}")
    dims = ", ".join(str(2 * l + 1) for l, _ in spec.outs)
    L.append(f"  // out dims: {dims}")

    # per-channel access to the [N, sum_a C*(2L_a+1)] hidden-irreps layout (block a at C*off_a + c*(2L_a+1) + M)
    for name, op in (("load_out", "o[{k}] = __ldg(row + (size_t)C * {off} + (size_t)c * {d} + {m});"),
                     ("store_out", "row[(size_t)C * {off} + (size_t)c * {d} + {m}] = o[{k}];")):
        const = "const " if name == "load_out" else ""
        oconst = "" if name == "load_out" else "const "
        L.append(f"  __device__ __forceinline__ static void {name}({const}float* __restrict__ row, int C, int c, {oconst}float (&o)[DOUT]) {{")
        off = 0
        for lo, _ in spec.outs:
            d = 2 * lo + 1
            for m in range(d):
                L.append("    " + op.format(k=off + m, off=off, d=d, m=m))
            off += d
        L.append("  }")

    # gather all monomials: per out index a, per nu: dict (k, M, idx) -> coef
    # (k indexes the reduced weight rows W' = B^T W, see so3.reduced_contraction)
    monos = [[so3.reduced_contraction(spec.coupling, o, nu)[0] for nu in range(1, corr + 1)] for o in spec.outs]

    # ---------------- forward
    L.append("  // out[M] per output irrep, concatenated: out[off_a + M]")
    L.append("  template <class WT> __device__ __forceinline__ static void fwd(const float (&x)[S], const WT& W, float (&out)[DOUT]) {")
    off = 0
    for a, (lo, _) in enumerate(spec.outs):
        for nu in range(1, corr + 1):
            byk = defaultdict(lambda: defaultdict(list))
            for (k, M, idx), c in monos[a][nu - 1].items():
                byk[k][M].append((idx, c))
            for k in sorted(byk):
                L.append("    {")
                L.append(f"      const float w = W.load({a}, {nu - 1}, {k});")
                for M in sorted(byk[k]):
                    expr = " + ".join(f"{_f(c)} * ({_mono_expr(idx)})" for idx, c in byk[k][M])
                    L.append(f"      out[{off + M}] = fmaf(w, {expr}, out[{off + M}]);")
                L.append("    }")
                L.append("    PML_SYM_SYNC")
        off += 2 * lo + 1
    L.append("  }")

    # ---------------- backward: dx += sum_M g_M d out_M / dx ; dW_k = sum_M g_M T_kM(x)
    L.append("  template <bool WANT_W, class WT, class GW> __device__ __forceinline__ static void bwd(const float (&x)[S], const WT& W, const float (&g)[DOUT], float (&dx)[S], GW& gw) {")
    off = 0
    for a, (lo, _) in enumerate(spec.outs):
        for nu in range(1, corr + 1):
            byk = defaultdict(lambda: defaultdict(list))
            for (k, M, idx), c in monos[a][nu - 1].items():
                byk[k][M].append((idx, c))
            for k in sorted(byk):
                L.append("    {")
                L.append(f"      const float w = W.load({a}, {nu - 1}, {k});")
                L.append("      if (WANT_W) {")
                parts = []
                for M in sorted(byk[k]):
                    expr = " + ".join(f"{_f(c)} * ({_mono_expr(idx)})" for idx, c in byk[k][M])
                    parts.append(f"g[{off + M}] * ({expr})")
                L.append(f"        gw.add({a}, {nu - 1}, {k}, {' + '.join(parts)});")
                L.append("      }")
                for M in sorted(byk[k]):
                    L.append(f"      const float s{M} = w * g[{off + M}];")
                # group derivative terms by j
                dj = defaultdict(list)
                for M in sorted(byk[k]):
                    for idx, c in byk[k][M]:
                        for j in sorted(set(idx)):
                            mult, rest = _dmono(idx, j)
                            dj[j].append((M, c * mult, rest))
                for j in sorted(dj):
                    expr = " + ".join(
                        f"{_f(c)} * (s{M}" + (f" * {_mono_expr(rest)}" if rest else "") + ")" for M, c, rest in dj[j]
                    )
                    L.append(f"      dx[{j}] += {expr};")
                L.append("    }")
                L.append("    PML_SYM_SYNC")
        off += 2 * lo + 1
    L.append("  }")

    # ---------------- directional second order: phi'(x; v) = d/de phi(x + e v)
    # dg_M  = sum_j v_j d out_M/dx_j                       (JVP)
    # ddx_j = sum_M g_M sum_i v_i d2 out_M / dx_i dx_j     (HVP)
    # ddW_k = sum_M g_M sum_i v_i d T_kM / dx_i
    L.append("  template <bool WANT_W, class WT, class GW> __device__ __forceinline__ static void dbl(const float (&x)[S], const float (&v)[S], const WT& W, const float (&g)[DOUT],")
    L.append("                                            float (&dg)[DOUT], float (&ddx)[S], GW& gw) {")
    off = 0
    for a, (lo, _) in enumerate(spec.outs):
        for nu in range(1, corr + 1):
            byk = defaultdict(lambda: defaultdict(list))
            for (k, M, idx), c in monos[a][nu - 1].items():
                byk[k][M].append((idx, c))
            for k in sorted(byk):
                L.append("    {")
                L.append(f"      const float w = W.load({a}, {nu - 1}, {k});")
                # JVP of basis: jv_M = sum_mono c * sum_i v_i d mono/dx_i
                for M in sorted(byk[k]):
                    parts = []
                    for idx, c in byk[k][M]:
                        for i in sorted(set(idx)):
                            mult, rest = _dmono(idx, i)
                            parts.append(f"{_f(c * mult)} * (v[{i}]" + (f" * {_mono_expr(rest)}" if rest else "") + ")")
                    L.append(f"      const float jv{M} = {' + '.join(parts)};")
                    L.append(f"      dg[{off + M}] = fmaf(w, jv{M}, dg[{off + M}]);")
                L.append("      if (WANT_W) {")
                L.append(f"        gw.add({a}, {nu - 1}, {k}, " + " + ".join(f"g[{off + M}] * jv{M}" for M in sorted(byk[k])) + ");")
                L.append("      }")
                if nu >= 2:
                    for M in sorted(byk[k]):
                        L.append(f"      const float s{M} = w * g[{off + M}];")
                    hj = defaultdict(list)
                    for M in sorted(byk[k]):
                        for idx, c in byk[k][M]:
                            for i in sorted(set(idx)):
                                m1, rest1 = _dmono(idx, i)
                                for j in sorted(set(rest1)):
                                    m2, rest2 = _dmono(rest1, j)
                                    hj[j].append((M, c * m1 * m2, i, rest2))
                    for j in sorted(hj):
                        expr = " + ".join(
                            f"{_f(c)} * (s{M} * v[{i}]" + (f" * {_mono_expr(rest)}" if rest else "") + ")"
                            for M, c, i, rest in hj[j]
                        )
                        L.append(f"      ddx[{j}] += {expr};")
                L.append("    }")
                L.append("    PML_SYM_SYNC")
        off += 2 * lo + 1
    L.append("  }")
    L.append("};")
    return "\n".join(L)


# ----------------------------------------------------------------------------------------------------------------------
def generate(out_dir: str = GEN_DIR) -> dict:
    os.makedirs(out_dir, exist_ok=True)
    tp_specs = default_tp_specs()
    sym_specs = default_sym_specs()
    registry = {"tp": {}, "sym": {}}

    parts = ["// GENERATED by physicsml_b200/codegen.py — do not edit.", "#pragma once",
             "using pml::f2mul; using pml::f2fma; using pml::f2dup; using pml::f2neg; using pml::red_add_f32x2;",
             "template <int ID> struct TPCode;", "template <int ID> struct TPGroups;",
             "template <int ID, int G> struct TPGroup;"]
    for ident, spec in enumerate(tp_specs):
        parts.append(_emit_tp(spec, ident))
        registry["tp"][spec.key] = {"id": ident, "np": spec.n_paths, "din": spec.d_in, "dout": spec.d_out, "nsh": spec.n_sh,
                                   "groups": len(tp_path_groups(spec))}
    _write_if_changed(os.path.join(out_dir, "tp_gen.cuh"), "\n".join(parts) + "\n")

    parts = ["// GENERATED by physicsml_b200/codegen.py — do not edit.", "#pragma once", "template <int ID> struct SymCode;"]
    for ident, spec in enumerate(sym_specs):
        parts.append(_emit_sym(spec, ident))
        registry["sym"][spec.key] = {"id": ident, "S": spec.S, "dout": spec.d_out, "K": spec.K, "K_full": spec.K_full, "outs": list(spec.outs)}
    _write_if_changed(os.path.join(out_dir, "sym_gen.cuh"), "\n".join(parts) + "\n")

    ids = [
        "// GENERATED by physicsml_b200/codegen.py — do not edit.",
        "#pragma once",
        f"#define PML_TP_NUM_SPECS {len(tp_specs)}",
        "#define PML_TP_FOREACH(X) " + " ".join(f"X({i})" for i in range(len(tp_specs))),
        f"#define PML_SYM_NUM_SPECS {len(sym_specs)}",
        "#define PML_SYM_FOREACH(X) " + " ".join(f"X({i})" for i in range(len(sym_specs))),
    ]
    _write_if_changed(os.path.join(out_dir, "pml_ids.h"), "\n".join(ids) + "\n")
    _write_if_changed(os.path.join(out_dir, "registry.json"), json.dumps(registry, indent=1, sort_keys=True) + "\n")
    return registry


def _write_if_changed(path: str, text: str) -> None:
    if os.path.exists(path):
        with open(path) as f:
            if f.read() == text:
                return
    with open(path, "w") as f:
        f.write(text)


if __name__ == "__main__":
    reg = generate()
    print(f"generated {len(reg['tp'])} tensor-product and {len(reg['sym'])} symmetric-contraction specialisations in {GEN_DIR}")
