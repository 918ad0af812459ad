"""Host-side plan objects: everything a kernel launch needs that depends only on the irreps configuration.

A plan is built once per module (fp64 on the host), owns the small device tables the C ABI wants and is referred to
from the ``torch.library`` ops by an integer handle (custom ops cannot take Python objects).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib, so3

_PLANS: list = []


def register(plan) -> int:
    _PLANS.append(plan)
    return len(_PLANS) - 1


def get(handle: int):
    return _PLANS[handle]


def _upload(structs, device) -> torch.Tensor:
    """ctypes array -> uint8 device tensor (synchronous copy; done once per plan and device)."""
    raw = bytes(structs)
    t = torch.frombuffer(bytearray(raw), dtype=torch.uint8)
    return t.to(device)


class LinearPlan:
    """o3.Linear semantics: y[n, o, w, m] = alpha_o * sum_{paths i->o} sum_u W[u, w] x[n, i, u, m],
    alpha_o = scale / sqrt(sum_{paths into o} mul_in).  Also covers nn.FullyConnectedNet layers (scalar irreps).

    out_layout "irreps": [N, sum mul*(2l+1)] ; "channel": [N, C, S] with S = sum (2l+1) over the output blocks
    (the reference's ``reshape_irreps``, irreps_tools.py:47-67, fused into the GEMM's output strides).
    """

    def __init__(self, irreps_in, irreps_out, scale: float = 1.0, out_layout: str = "irreps",
                 in_layout: str = "irreps") -> None:
        # in_layout "component": every input block is stored component-major, x[n, off_b + m*mul + u] (what
        # tp_scatter writes with TPPlan(out_layout="component")): unit stride along the reduction index u
        if in_layout not in ("irreps", "component"):
            raise ValueError(in_layout)
        self.in_layout = in_layout
        mmaj = in_layout == "component"
        self.irreps_in = so3.parse_irreps(irreps_in)
        self.irreps_out = so3.parse_irreps(irreps_out)
        self.scale = float(scale)
        self.out_layout = out_layout
        self.d_in = so3.irreps_dim(self.irreps_in)
        self.d_out = so3.irreps_dim(self.irreps_out)
        in_off = so3.irreps_offsets(self.irreps_in)
        paths = []
        w_off = 0
        for i, (mi, li, pi) in enumerate(self.irreps_in):
            for o, (mo, lo, po) in enumerate(self.irreps_out):
                if (li, pi) == (lo, po):
                    paths.append((i, o, w_off))
                    w_off += mi * mo
        self.paths = paths
        self.weight_numel = w_off
        fan_in = [0] * len(self.irreps_out)
        for i, o, _ in paths:
            fan_in[o] += self.irreps_in[i][0]
        self.alpha = [self.scale / math.sqrt(f) if f > 0 else 0.0 for f in fan_in]
        # output addressing
        if out_layout == "irreps":
            o_off = so3.irreps_offsets(self.irreps_out)
            o_cs = [2 * l + 1 for _, l, _ in self.irreps_out]
            self.out_row = self.d_out
        elif out_layout == "channel":
            muls = {m for m, _, _ in self.irreps_out}
            if len(muls) != 1:
                raise ValueError("channel layout needs equal multiplicities")
            S = sum(2 * l + 1 for _, l, _ in self.irreps_out)
            o_off, acc = [], 0
            for _, l, _ in self.irreps_out:
                o_off.append(acc)
                acc += 2 * l + 1
            o_cs = [S] * len(self.irreps_out)
            self.out_row = self.d_out
            self.channels, self.S = muls.pop(), S
        else:
            raise ValueError(out_layout)
        self.has_unfed_output = any(f == 0 for f in fan_in)
        # scalar layer with a short reduction and a wide output (the output layer of the radial MLP): (K, N, alpha) for
        # the streaming kernels of csrc/gemm_wide.cu, else None
        self.wide = None
        self.narrow = None
        if (len(self.irreps_in) == 1 and len(self.irreps_out) == 1 and self.irreps_in[0][1] == 0
                and self.irreps_out[0][1] == 0 and len(paths) == 1 and out_layout == "irreps"):
            k_, n_ = self.irreps_in[0][0], self.irreps_out[0][0]
            if k_ <= 64 and k_ % 4 == 0 and n_ % 64 == 0 and n_ >= 256:
                self.wide = (k_, n_, self.alpha[0])
            # narrow scalar layers (the hidden layers of the radial MLP, 64 -> 64): both y = x W and dx = g W^T are
            # "reduce over <= a few 32-chunks into <= 64 outputs", which is what the streaming dx kernel computes
            self.narrow = None
            if k_ % 32 == 0 and n_ % 32 == 0 and k_ <= 64 and n_ <= 64:
                self.narrow = (k_, n_, self.alpha[0])
        self.max_mul_in = max(m for m, _, _ in self.irreps_in)
        self.max_mul_out = max(m for m, _, _ in self.irreps_out)
        self.max_d = max(2 * l + 1 for _, l, _ in self.irreps_out)

        def prob():
            p = _lib.GemmProblem()
            for c in range(_lib.GEMM_MAX_CHUNKS):
                p.cscale[c] = 1.0
            return p

        # ---- forward: one problem per fed output block; chunks = its input blocks
        fwd = []
        for o, (mo, lo, _) in enumerate(self.irreps_out):
            ch = [(i, w) for i, oo, w in paths if oo == o]
            if not ch:
                continue
            d = 2 * lo + 1
            for c0 in range(0, len(ch), _lib.GEMM_MAX_CHUNKS):
                sub = ch[c0 : c0 + _lib.GEMM_MAX_CHUNKS]
                p = prob()
                p.nchunks = len(sub)
                for c, (i, w) in enumerate(sub):
                    p.a_off[c], p.b_off[c], p.kc[c] = in_off[i], w, self.irreps_in[i][0]
                p.c_off = o_off[o]
                p.a_rs, p.a_cs, p.a_bs = self.d_in, d, 1
                if mmaj:
                    muls_in = {self.irreps_in[i][0] for i, _ in sub}
                    if len(muls_in) != 1:
                        raise NotImplementedError("component-major input blocks of one irrep with different multiplicities")
                    p.a_cs, p.a_bs = 1, muls_in.pop()
                p.b_rs, p.b_cs, p.b_bs = mo, 1, 0
                p.c_rs, p.c_cs, p.c_bs = self.out_row, o_cs[o], 1
                p.M, p.N, p.nbatch, p.alpha = -1, mo, d, self.alpha[o]
                p.accumulate = 1 if c0 > 0 else 0
                fwd.append(p)
        # problems that accumulate onto an earlier problem's C would race inside one launch
        self.fwd_serial = any(p.accumulate for p in fwd)
        self._fwd = fwd
        # ---- dx: one problem per fed input block; chunks = the output blocks it feeds (per-chunk alpha)
        bx = []
        for i, (mi, li, _) in enumerate(self.irreps_in):
            ch = [(o, w) for ii, o, w in paths if ii == i]
            if not ch:
                continue
            if len(ch) > _lib.GEMM_MAX_CHUNKS:
                raise NotImplementedError("an input block feeding more than 8 output blocks")
            d = 2 * li + 1
            p = prob()
            p.nchunks = len(ch)
            for c, (o, w) in enumerate(ch):
                p.a_off[c], p.b_off[c], p.kc[c], p.cscale[c] = o_off[o], w, self.irreps_out[o][0], self.alpha[o]
            p.c_off = in_off[i]
            # A = g (same addressing as the forward's C); all chunks of one input block share the output irrep's strides
            o0 = ch[0][0]
            p.a_rs, p.a_cs, p.a_bs = self.out_row, o_cs[o0], 1
            mo = self.irreps_out[o0][0]
            if any(self.irreps_out[o][0] != mo for o, _ in ch):
                raise NotImplementedError("output blocks of one irrep with different multiplicities")
            p.b_rs, p.b_cs, p.b_bs = 1, mo, 0  # B[k = w, col = u] = W[u, w]
            p.c_rs, p.c_cs, p.c_bs = (self.d_in, 1, mi) if mmaj else (self.d_in, d, 1)
            p.M, p.N, p.nbatch, p.alpha, p.accumulate = -1, mi, d, 1.0, 0
            bx.append(p)
        self.has_unfed_input = len(bx) < len(self.irreps_in)
        # longest reduction of any problem (selects the accumulator-promoting tensor-core kernel above 128)
        self.max_k_fwd = max([sum(p.kc[c] for c in range(p.nchunks)) for p in fwd] + [1])
        self.max_k_bwd_x = max([sum(p.kc[c] for c in range(p.nchunks)) for p in bx] + [1])
        self._bwd_x = bx
        # ---- dW: one problem per path, K = nodes (runtime), the (2l+1) batches reduce into the same C (atomic epilogue)
        bw = []
        for i, o, w in paths:
            mi, li, _ = self.irreps_in[i]
            mo = self.irreps_out[o][0]
            d = 2 * li + 1
            p = prob()
            p.nchunks = 1
            p.a_off[0], p.b_off[0], p.kc[0] = in_off[i], o_off[o], -1
            p.c_off = w
            p.a_rs, p.a_cs, p.a_bs = d, self.d_in, 1  # A[u, k = n] = x[n, in_off + u d + m]
            if mmaj:
                p.a_rs, p.a_bs = 1, mi  # x[n, in_off + m mul + u]
            p.b_rs, p.b_cs, p.b_bs = self.out_row, o_cs[o], 1  # B[k = n, w] = g[n, o_off + w cs + m]
            p.c_rs, p.c_cs, p.c_bs = mo, 1, 0
            p.M, p.N, p.nbatch, p.alpha, p.accumulate = mi, mo, d, self.alpha[o], 0
            bw.append(p)
        self._bwd_w = bw
        self.bwd_w_batches = sum(int(p.nbatch) for p in bw)  # (problem, batch) pairs that exist (the grid has max_d per problem)
        # ---- weight gradients of "few inputs -> many outputs" layers (the radial MLP's output layer: 64 -> P*C): computed
        # TRANSPOSED, dW^T[w, u] = sum_n g[n, w] x[n, u], so that the wide dimension fills the 128-row MMA tile and the narrow
        # one the 32/64-column tile.  As dW[u, w] a 64-row A tile is padded to 128 rows whose staging costs the same
        # instructions as real rows: half the k-blocks for the same work (config 5, 64 -> 4352: 68 column tiles x 4304
        # k-blocks -> 34 x 4304).  ops.irrep_linear_bwd_w passes (g, x) as (A, B) for these tables.
        self._bwd_w_t = None
        if bw and all(self.irreps_in[i][0] <= 64 and self.irreps_out[o][0] >= 128 for i, o, _ in paths):
            self._bwd_w_t = []
            for p0, (i, o, w_) in zip(bw, paths):
                mi, mo = self.irreps_in[i][0], self.irreps_out[o][0]
                p = prob()
                p.nchunks = 1
                p.a_off[0], p.b_off[0], p.kc[0] = p0.b_off[0], p0.a_off[0], -1
                p.c_off = w_
                p.a_rs, p.a_cs, p.a_bs = p0.b_cs, p0.b_rs, p0.b_bs  # A'[w, k = n] = g[n, o_off + w cs + m]
                p.b_rs, p.b_cs, p.b_bs = p0.a_cs, p0.a_rs, p0.a_bs  # B'[k = n, u] = x[n, in_off + u d + m]
                p.c_rs, p.c_cs, p.c_bs = 1, mo, 0                   # C'[w, u] = dW[u, w]
                p.M, p.N, p.nbatch, p.alpha, p.accumulate = mo, mi, p0.nbatch, p0.alpha, 0
                self._bwd_w_t.append(p)
        # ---- component-major operand copies (ops.irrep_linear / irrep_linear_bwd_x on node- or edge-sized inputs): the
        # reduction index u (forward) / w (cotangent) has stride (2l+1) in the e3nn order and S in the channel layout —
        # one 32-byte sector per element for the GEMM's operand loads.  `colmap_in` / `colmap_out` re-order a row to
        # m*mul + u inside every block (pml_permute_columns, a few microseconds) and `_fwd_c` / `_bwd_x_c` are the same
        # problems addressed in that copy.
        import copy as _copy

        self.in_strided = (not mmaj) and any(l > 0 and m_ >= 16 for m_, l, _ in self.irreps_in)
        self.out_strided = any((cs != 1) and self.irreps_out[o][0] >= 16 for o, cs in enumerate(o_cs))
        oc_off = so3.irreps_offsets(self.irreps_out)
        cm_in = np.zeros(self.d_in, dtype=np.int32)
        for i, (mi, li, _) in enumerate(self.irreps_in):
            d = 2 * li + 1
            for m_ in range(d):
                for u in range(mi):
                    cm_in[in_off[i] + m_ * mi + u] = in_off[i] + u * d + m_
        cm_out = np.zeros(self.d_out, dtype=np.int32)
        for o, (mo, lo, _) in enumerate(self.irreps_out):
            d = 2 * lo + 1
            for m_ in range(d):
                for w in range(mo):
                    cm_out[oc_off[o] + m_ * mo + w] = o_off[o] + w * o_cs[o] + m_
        self._colmap_in, self._colmap_out = cm_in, cm_out
        self._fwd_c = None
        if self.in_strided and not self.fwd_serial:
            self._fwd_c = []
            for p0 in fwd:
                p = _copy.copy(p0)
                muls_in = {int(p.kc[c]) for c in range(p.nchunks)}
                if len(muls_in) != 1:
                    self._fwd_c = None
                    break
                p.a_cs, p.a_bs = 1, muls_in.pop()
                self._fwd_c.append(p)
        self._bwd_x_c = None
        if self.out_strided:
            self._bwd_x_c = []
            for p0 in bx:
                p = _copy.copy(p0)
                mo = int(p.kc[0])
                for c in range(p.nchunks):  # chunk c = output block o fed by this input block: find it by its offset
                    o = o_off.index(int(p.a_off[c]))
                    p.a_off[c] = oc_off[o]
                p.a_cs, p.a_bs = 1, mo
                self._bwd_x_c.append(p)
        self._dev: dict = {}
        self._dev_c: dict = {}
        self.handle = register(self)

    def tables_c(self, device):
        """(colmap_in, fwd table on the component-major x) , (colmap_out, dx table on the component-major g); None where the
        operand already has unit stride"""
        key = str(device)
        if key not in self._dev_c:
            def up(lst):
                return None if not lst else _upload((_lib.GemmProblem * len(lst))(*lst), device)

            f = (torch.tensor(self._colmap_in, device=device), up(self._fwd_c)) if self._fwd_c else None
            b = (torch.tensor(self._colmap_out, device=device), up(self._bwd_x_c)) if self._bwd_x_c else None
            self._dev_c[key] = (f, b)
        return self._dev_c[key]

    def tables(self, device):
        key = str(device)
        if key not in self._dev:
            def up(lst):
                if not lst:
                    return None
                return _upload((_lib.GemmProblem * len(lst))(*lst), device)

            self._dev[key] = (up(self._fwd), up(self._bwd_x), up(self._bwd_w))
            self._dev[key + ":t"] = up(self._bwd_w_t) if self._bwd_w_t else None
        return self._dev[key]

    def table_bwd_w_t(self, device):
        self.tables(device)
        return self._dev[str(device) + ":t"]


class ElementLinearPlan:
    """FullyConnectedTensorProduct(x, Zx0e attrs) semantics."""

    def __init__(self, irreps_in, num_elements: int, irreps_out, out_layout: str = "irreps") -> None:
        self.irreps_in = so3.parse_irreps(irreps_in)
        self.irreps_out = so3.parse_irreps(irreps_out)
        self.Z = int(num_elements)
        self.d_in = so3.irreps_dim(self.irreps_in)
        self.d_out = so3.irreps_dim(self.irreps_out)
        self.out_layout = out_layout
        in_off = so3.irreps_offsets(self.irreps_in)
        if out_layout == "irreps":
            o_off = so3.irreps_offsets(self.irreps_out)
            o_ws = [2 * l + 1 for _, l, _ in self.irreps_out]
        else:
            S = sum(2 * l + 1 for _, l, _ in self.irreps_out)
            o_off, acc = [], 0
            for _, l, _ in self.irreps_out:
                o_off.append(acc)
                acc += 2 * l + 1
            o_ws = [S] * len(self.irreps_out)
        paths, w_off = [], 0
        for i, (mi, li, pi) in enumerate(self.irreps_in):
            for o, (mo, lo, po) in enumerate(self.irreps_out):
                if (li, pi) == (lo, po):
                    paths.append((i, o, w_off))
                    w_off += mi * self.Z * mo
        self.weight_numel = w_off
        fan = [0] * len(self.irreps_out)
        for i, o, _ in paths:
            fan[o] += self.irreps_in[i][0] * self.Z
        if len(paths) > _lib.EL_MAX_PATHS:
            raise NotImplementedError("more than 8 element-linear paths")
        outs = [o for _, o, _ in paths]
        if len(set(outs)) != len(outs):
            raise NotImplementedError("several input blocks feeding one output block")
        self.has_unfed_output = len(set(outs)) < len(self.irreps_out)
        self.has_unfed_input = len({i for i, _, _ in paths}) < len(self.irreps_in)
        p = _lib.ElPlan()
        p.npaths = len(paths)
        for k, (i, o, w) in enumerate(paths):
            mi, li, _ = self.irreps_in[i]
            p.d[k], p.mul_in[k], p.mul_out[k] = 2 * li + 1, mi, self.irreps_out[o][0]
            p.w_off[k], p.x_off[k], p.o_off[k], p.o_ws[k] = w, in_off[i], o_off[o], o_ws[o]
            p.alpha[k] = 1.0 / math.sqrt(fan[o])
        p.x_rs, p.o_rs = self.d_in, self.d_out
        self.c_plan = p
        self.n_launch = len({2 * self.irreps_in[i][1] + 1 for i, _, _ in paths})
        self._build_segment_templates(paths, in_off, o_off, o_ws)
        self._seg_dev: dict = {}
        self.handle = register(self)

    def _build_segment_templates(self, paths, in_off, o_off, o_ws) -> None:
        """Tensor-core route (ops.element_linear with a SpeciesPlan): on nodes GROUPED by element the op is one dense GEMM
        per (path, element) over a contiguous row range.  Operands are compact component-major copies
            xs[pos, cin_i + m*mul_in + u] = x[order[pos], in_off_i + u*d + m]      (fed input blocks only)
            ys[pos, cout_o + m*mul_out + w] <-> out[order[pos], o_off_o + w*o_ws + m]  (fed output blocks only)
        so that the reduction index and the output columns have unit stride.  The templates below are completed on the
        device by pml_segment_problems (group start / size of every element): problem p*Z + v of each table."""
        Z = self.Z
        fed_in = sorted({i for i, _, _ in paths})
        fed_out = sorted({o for _, o, _ in paths})
        cin, cout, acc = {}, {}, 0
        for i in fed_in:
            cin[i] = acc
            acc += self.irreps_in[i][0] * (2 * self.irreps_in[i][1] + 1)
        self.d_in_c = acc
        acc = 0
        for o in fed_out:
            cout[o] = acc
            acc += self.irreps_out[o][0] * (2 * self.irreps_out[o][1] + 1)
        self.d_out_c = acc
        cm_in = np.zeros(max(self.d_in_c, 1), dtype=np.int32)
        for i in fed_in:
            mi, li, _ = self.irreps_in[i]
            d = 2 * li + 1
            for m_ in range(d):
                for u in range(mi):
                    cm_in[cin[i] + m_ * mi + u] = in_off[i] + u * d + m_
        cm_out = np.zeros(max(self.d_out_c, 1), dtype=np.int32)
        for o in fed_out:
            mo, lo, _ = self.irreps_out[o]
            d = 2 * lo + 1
            for m_ in range(d):
                for w in range(mo):
                    cm_out[cout[o] + m_ * mo + w] = o_off[o] + w * o_ws[o] + m_
        self._cm_in, self._cm_out = cm_in, cm_out

        def prob():
            q = _lib.GemmProblem()
            for c in range(_lib.GEMM_MAX_CHUNKS):
                q.cscale[c] = 1.0
            q.nchunks, q.accumulate = 1, 0
            return q

        fwd, bx, bw, mf, mx, mw = [], [], [], [], [], []
        for k, (i, o, w_off) in enumerate(paths):
            mi, li, _ = self.irreps_in[i]
            mo = self.irreps_out[o][0]
            d = 2 * li + 1
            alpha = float(self.c_plan.alpha[k])
            for v in range(Z):
                wv = w_off + v * mo  # W[u, v, w] at w_off + (u*Z + v)*mo + w
                q = prob()  # ys = alpha * xs W_v over the rows of element v
                q.a_off[0], q.b_off[0], q.kc[0], q.c_off = cin[i], wv, mi, cout[o]
                q.a_rs, q.a_cs, q.a_bs = self.d_in_c, 1, mi
                q.b_rs, q.b_cs, q.b_bs = Z * mo, 1, 0
                q.c_rs, q.c_cs, q.c_bs = self.d_out_c, 1, mo
                q.M, q.N, q.nbatch, q.alpha = 0, mo, d, alpha
                fwd.append(q)
                mf.append(_lib.SegMeta(v, 0, self.d_in_c, 0, self.d_out_c))
                q = prob()  # dxs = alpha * gs W_v^T
                q.a_off[0], q.b_off[0], q.kc[0], q.c_off = cout[o], wv, mo, cin[i]
                q.a_rs, q.a_cs, q.a_bs = self.d_out_c, 1, mo
                q.b_rs, q.b_cs, q.b_bs = 1, Z * mo, 0  # B[k = w, col = u] = W[u, v, w]
                q.c_rs, q.c_cs, q.c_bs = self.d_in_c, 1, mi
                q.M, q.N, q.nbatch, q.alpha = 0, mi, d, alpha
                bx.append(q)
                mx.append(_lib.SegMeta(v, 0, self.d_out_c, 0, self.d_in_c))
                q = prob()  # dW[:, v, :] = alpha * sum over the element's rows and the 2l+1 components of xs^T gs
                q.a_off[0], q.b_off[0], q.kc[0], q.c_off = cin[i], cout[o], 0, wv
                q.a_rs, q.a_cs, q.a_bs = 1, self.d_in_c, mi     # A[u, k = pos]
                q.b_rs, q.b_cs, q.b_bs = self.d_out_c, 1, mo    # B[k = pos, w]
                q.c_rs, q.c_cs, q.c_bs = Z * mo, 1, 0
                q.M, q.N, q.nbatch, q.alpha = mi, mo, d, alpha
                bw.append(q)
                mw.append(_lib.SegMeta(v, 1, self.d_in_c, self.d_out_c, 0))
        self._seg_tmpl = fwd + bx + bw
        self._seg_meta = mf + mx + mw
        self.seg_nprob = len(fwd)
        self.seg_max_mi = max([self.irreps_in[i][0] for i, _, _ in paths] + [1])
        self.seg_max_mo = max([self.irreps_out[o][0] for _, o, _ in paths] + [1])
        self.seg_max_d = max([2 * self.irreps_in[i][1] + 1 for i, _, _ in paths] + [1])
        # worth three extra launches (two permutes + the table kernel) only when the contraction is sizeable: the scalar
        # residual layers (one 128 -> 128 block) stay on the FFMA kernels
        work = sum(self.irreps_in[i][0] * self.irreps_out[o][0] * (2 * self.irreps_in[i][1] + 1) for i, o, _ in paths)
        self.seg_ok = bool(paths) and work >= 4 * 64 * 64 and self.seg_max_mi >= 16 and self.seg_max_mo >= 16

    def segment_tables(self, device):
        """(templates [3 * seg_nprob problems: forward, dx, dW], metas, colmap_in, colmap_out) on `device`"""
        key = str(device)
        if key not in self._seg_dev:
            n = len(self._seg_tmpl)
            self._seg_dev[key] = (
                _upload((_lib.GemmProblem * n)(*self._seg_tmpl), device),
                _upload((_lib.SegMeta * n)(*self._seg_meta), device),
                torch.tensor(self._cm_in, device=device),
                torch.tensor(self._cm_out, device=device),
            )
        return self._seg_dev[key]


class TPPlan:
    def __init__(self, node_irreps, sh_lmax: int, target_irreps, sh_parity: int = -1, out_layout: str = "irreps") -> None:
        # out_layout "component": every output block is written component-major, out[n, C*off_b + m*C + u], instead of
        # the e3nn order u*(2l+1) + m: packed 8-byte stores from the kernel and a unit-stride reduction index for the
        # linear that follows (LinearPlan(in_layout="component")).  Internal to a module: never leaves it.
        if out_layout not in ("irreps", "component"):
            raise ValueError(out_layout)
        self.out_layout_id = 1 if out_layout == "component" else 0
        self.spec = so3.build_tp_spec(node_irreps, sh_lmax, target_irreps, sh_parity)
        muls = {m for m, _, _ in so3.parse_irreps(node_irreps)}
        self.C = muls.pop()
        reg = _lib.registry()["tp"]
        if self.spec.key not in reg:
            raise NotImplementedError(
                f"no compiled tensor-product specialisation for {self.spec.key}; add it to "
                "physicsml_b200.codegen.default_tp_specs() and rebuild"
            )
        self.spec_id = reg[self.spec.key]["id"]
        self.n_groups = int(reg[self.spec.key].get("groups", 1))  # path groups of the pipelined kernels (one launch each)
        self.NP, self.DIN, self.DOUT, self.NSH = self.spec.n_paths, self.spec.d_in, self.spec.d_out, self.spec.n_sh
        # irreps of the aggregated message, sorted, one block per path (mul C each)
        self.out_irreps = [(self.C, l, p) for l, p in self.spec.out_ls]
        self.handle = register(self)


class SymPlan:
    def __init__(self, lmax: int, outs, corr: int, channels: int, num_elements: int) -> None:
        from .codegen import SymSpec

        self.spec = SymSpec(lmax, outs, corr)
        reg = _lib.registry()["sym"]
        if self.spec.key not in reg:
            raise NotImplementedError(
                f"no compiled symmetric-contraction specialisation for {self.spec.key}; add it to "
                "physicsml_b200.codegen.default_sym_specs() and rebuild"
            )
        self.spec_id = reg[self.spec.key]["id"]
        self.K = self.spec.K  # [out][nu-1], rows of the reduced weights the kernels read
        self.K_full = self.spec.K_full  # rows of the reference's parameters
        self._B = {}  # device -> list of [K_full, K] matrices (None where the reduction is the identity)
        self.S, self.DOUT, self.C, self.Z, self.corr = self.spec.S, self.spec.d_out, channels, num_elements, corr
        self.n_out = len(self.spec.outs)
        self.handle = register(self)

    def reduce_weights(self, weights):
        """reference parameters [Z, K_full, C] (flat list ordered (out a, nu)) -> W' = B^T W [Z, K, C]: the polynomials
        the K_full rows multiply are linearly dependent (so3.reduced_contraction); plain differentiable torch ops, so
        gradients reach the parameters through autograd."""
        dev = weights[0].device
        if dev not in self._B:
            mats = []
            for a, o in enumerate(self.spec.outs):
                for nu in range(1, self.corr + 1):
                    B = so3.reduced_contraction(self.spec.coupling, o, nu)[1]
                    ident = B.shape[0] == B.shape[1] and np.array_equal(B, np.eye(B.shape[0]))
                    mats.append(None if ident else torch.tensor(B.T.copy(), dtype=torch.float32, device=dev))
            self._B[dev] = mats
        return [w if Bt is None else torch.matmul(Bt, w) for w, Bt in zip(weights, self._B[dev])]

    def weight_struct(self, weights, grads=None):
        """weights: flat list ordered (out a, nu = 1..corr)."""
        st = _lib.SymWeights()
        k = 0
        for a in range(self.n_out):
            for nu in range(self.corr):
                st.w[a][nu] = weights[k].data_ptr()
                st.gw[a][nu] = grads[k].data_ptr() if grads is not None else None
                st.K[a][nu] = self.K[a][nu]
                k += 1
        return st


SILU_2MOM = 1.6791767923989418  # e3nn normalize2mom constants
TANH_2MOM = 1.5937334472592692


class GatePlan:
    """NequIP ``Gate`` (reference models/nequip/modules/_gate.py:88-152 as configured in convnet_layer.py:44-77):
    input row = sort(scalars + gates + gated).simplify(), output row = scalars + gated.  SiLU on even scalars / gates,
    tanh on odd ones, each scaled by its normalize2mom constant; gated block b (mul, l) is multiplied channelwise by
    the activated gate scalars b."""

    def __init__(self, irreps_scalars, irreps_gates, irreps_gated) -> None:
        sc = so3.parse_irreps(irreps_scalars)
        gt = so3.parse_irreps(irreps_gates)
        gd = so3.parse_irreps(irreps_gated)
        if any(l != 0 for _, l, _ in sc + gt):
            raise ValueError("gate scalars / gates must be l = 0")
        if any(p != 1 for _, _, p in gt):
            raise NotImplementedError("odd (0o) gates flip the parity of the gated irreps; not supported")
        if [m for m, _, _ in gt] != [m for m, _, _ in gd]:
            raise ValueError("one gate scalar per gated channel is required")
        # the reference's _Sortcut: concatenate, stable-sort by (l, p), remember where every piece went
        pieces = [("s", i, ir) for i, ir in enumerate(sc)] + [("g", i, ir) for i, ir in enumerate(gt)] + [
            ("x", i, ir) for i, ir in enumerate(gd)]
        srt, _ = so3.sort_irreps([ir for _, _, ir in pieces])
        order = sorted(range(len(pieces)), key=lambda i: ((pieces[i][2][1], pieces[i][2][2]), i))
        off, where = 0, {}
        for idx in order:
            kind, i, (m, l, _p) = pieces[idx]
            where[(kind, i)] = off
            off += m * (2 * l + 1)
        self.irreps_in = so3.simplify_irreps(srt)
        self.irreps_out = so3.simplify_irreps(sc + gd)
        self.d_in, self.d_out = off, so3.irreps_dim(sc + gd)
        st = _lib.GatePlanStruct()
        if len(sc) > _lib.GATE_MAX_SEG or len(gd) > _lib.GATE_MAX_SEG:
            raise NotImplementedError("too many gate segments")
        out_off, items = 0, 0
        st.n_scal = len(sc)
        for i, (m, _l, p) in enumerate(sc):
            st.s_in[i], st.s_out[i], st.s_mul[i] = where[("s", i)], out_off, m
            st.s_kind[i], st.s_c[i] = (0, SILU_2MOM) if p == 1 else (1, TANH_2MOM)
            out_off += m
            items += m
        st.n_gated = len(gd)
        for i, (m, l, _p) in enumerate(gd):
            st.g_x[i], st.g_gate[i], st.g_out[i], st.g_mul[i], st.g_d[i] = where[("x", i)], where[("g", i)], out_off, m, 2 * l + 1
            st.g_kind[i], st.g_c[i] = 0, SILU_2MOM
            out_off += m * (2 * l + 1)
            items += m
        st.d_in, st.d_out, st.items = self.d_in, self.d_out, items
        self.c_plan = st
        self.handle = register(self)
