"""configs[3] of BASELINE.json: MACE-medium energy+forces inference on a periodic water box (24k atoms, r_cut 6 A, full
PBC): neighbour list on the GPU, fused edge kernels at HBM-bound size.  Reports the per-kernel roofline of the fused
edge kernels (SURVEY §8d algorithmic bytes), the neighbour-list time and the whole evaluation.
usage: python scripts/bench_large.py [n_molecules] [reps]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from physicsml_b200 import ops, plans  # noqa: E402
from physicsml_b200 import workloads as wl  # noqa: E402
from physicsml_b200.mace import PooledMACEModule, compute_forces_by_gradient  # noqa: E402
from physicsml_b200.neighbour_list import radius_graph  # noqa: E402

n_mol = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
box = 62.14 * (n_mol / 8000.0) ** (1.0 / 3.0)
dev = torch.device("cuda:0")
pos, z, cell = wl.water_box(n_mol, box)
rc = 6.0
kw = dict(cut_off=rc, num_bessel=8, num_polynomial_cutoff=5, max_ell=3, num_interactions=2, num_node_feats=10, num_edge_feats=0,
          hidden_irreps="128x0e + 128x1o", avg_num_neighbours=90.0, correlation=3)
torch.manual_seed(0)
model = PooledMACEModule(mlp_irreps="16x0e", **kw).to(dev)
pos_d, cell_d = pos.to(dev), cell.to(dev)
attrs = torch.nn.functional.one_hot(z, 10).float().to(dev)
N = pos.shape[0]


def ev(t0, t1):
    return t0.elapsed_time(t1)


def nl():
    return radius_graph(pos_d, rc, pbc=(True, True, True), cell=cell_d)


for _ in range(2):
    ei, sh, plan = nl()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(reps):
    ei, sh, plan = nl()
torch.cuda.synchronize()
nl_ms = (time.perf_counter() - t0) / reps * 1e3
E = ei.shape[1]
print(f"atoms {N} edges {E} mean degree {E / N:.1f} | GPU neighbour list {nl_ms:.2f} ms (host-timed, includes the E readback)")


def evaluate():
    d = {"coordinates": pos_d.clone().requires_grad_(True), "edge_index": ei, "cell": cell_d, "cell_shift_vector": sh,
         "node_attrs": attrs, "batch": torch.zeros(N, dtype=torch.long, device=dev), "num_graphs": torch.tensor(1), "_edge_plan": plan}
    e = model(d)["y_graph_scalars"]
    f = compute_forces_by_gradient(e, d["coordinates"])
    return e, f


for _ in range(2):
    evaluate()
torch.cuda.synchronize()
print(f"peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
ops.enable_kernel_timing(["pml_tp_fwd", "pml_tp_bwd", "pml_sym_fwd", "pml_sym_bwd", "pml_gemm_grouped", "pml_gemm_grouped_tc",
                          "pml_edge_featurise_fwd", "pml_edge_featurise_bwd", "pml_element_linear_fwd", "pml_element_linear_bwd_x",
                          "pml_wide_fwd", "pml_wide_bwd_x", "pml_act"])
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    evaluate()
e1.record()
torch.cuda.synchronize()
ms = ev(e0, e1) / reps
print(f"energy+forces evaluation {ms:.2f} ms  -> {N / ms * 1e3:.3e} atoms/s (edges precomputed); with NL {N / (ms + nl_ms) * 1e3:.3e} atoms/s")
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
rows = []
for name, recs in ops.KERNEL_TIMING.items():
    by = {}
    for tag, a, b in recs:
        by.setdefault(tag if name.startswith("pml_tp") or name.startswith("pml_sym") else 0, []).append(a.elapsed_time(b))
    for tag, ts in by.items():
        tot = sum(ts) / reps
        line = f"{name:28s} spec/tag {tag!s:>14s}  launches/eval {len(ts) / reps:5.1f}  ms/eval {tot:8.3f}  avg {sum(ts) / len(ts):8.3f} ms"
        if name in ("pml_tp_fwd", "pml_tp_bwd"):
            spec = tag[0] if isinstance(tag, tuple) else tag  # cotangent launches are tagged (spec, need_h, need_y, need_r)
            pl = next(p for p in plans._PLANS if getattr(p, "spec_id", None) == spec and hasattr(p, "DIN"))
            fwd_b = E * (4 * pl.NP * pl.C + 4 * pl.NSH + 16) + N * 4 * (pl.C * pl.DIN + pl.C * pl.DOUT)
            if name == "pml_tp_bwd":  # reads R,Y,h,g ; writes dR, dY (+ dh atomics): SURVEY §8d cotangent formulas combined
                fwd_b = E * (8 * pl.NP * pl.C + 8 * pl.NSH + 16) + N * 4 * (2 * pl.C * pl.DIN + pl.C * pl.DOUT)
            gbs = fwd_b / (sum(ts) / len(ts) * 1e-3) / 1e9
            line += f"  | algorithmic {fwd_b / 1e9:6.2f} GB -> {gbs:7.1f} GB/s = {100 * gbs / peak:5.1f}% of measured HBM peak"
        rows.append((tot, line))
for _, line in sorted(rows, reverse=True):
    print(line)

if len(sys.argv) > 3 and sys.argv[3] == "shapes":  # per-shape time of the general grouped GEMM (event-bracketed)
    import collections

    ops.KERNEL_TIMING.clear()
    records, orig = [], ops._gemm

    def wrapped(gname, a, b, c, table, nprob, m_runtime, max_m, max_n, max_batch, splitk, max_k):
        x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        x0.record()
        orig(gname, a, b, c, table, nprob, m_runtime, max_m, max_n, max_batch, splitk, max_k)
        x1.record()
        records.append(((sys._getframe(1).f_code.co_name, nprob, m_runtime, max_m, max_n, max_batch, splitk, max_k), x0, x1))

    ops._gemm = wrapped
    evaluate()
    torch.cuda.synchronize()
    agg = collections.OrderedDict()
    for key, x0, x1 in records:
        v = agg.setdefault(key, [0, 0.0])
        v[0] += 1
        v[1] += x0.elapsed_time(x1)
    print("general GEMM by (caller, nprob, M_runtime, max_M, max_N, batch, splitk, max_k): launches, ms each, ms total")
    for key, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{str(key):90s} {n:3d} {t / n:8.3f} {t:8.3f}")
