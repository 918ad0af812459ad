"""Microbenchmark of the fused tensor-product + scatter kernels (K4) at the edge counts of BASELINE configs 2 and 4:
TMA-pipelined persistent kernels vs the register-streaming kernels, CUDA-event timed, L2 flushed between launches,
algorithmic bytes from SURVEY §8d.   usage: python scripts/bench_tp.py [reps] [case-substring] [modes, e.g. 1,0,6]
(modes = pml_tp_set_simple bits: 1 streaming kernels, 2 / 4 flip the producer placement of fwd / bwd)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from physicsml_b200 import _lib, ops  # noqa: E402
from physicsml_b200.plans import TPPlan  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
only = sys.argv[2] if len(sys.argv) > 2 else ""
modes = [int(m) for m in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 0, 6]
dev = torch.device("cuda:0")
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
peak = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

CASES = [  # name, node irreps, lmax, C, N, mean degree
    ("cfg2 L0", "128x0e", 3, 128, 1304, 32),
    ("cfg2 L1", "128x0e+128x1o", 3, 128, 1304, 32),
    ("cfg4 L0", "128x0e", 3, 128, 24000, 85),
    ("cfg4 L1", "128x0e+128x1o", 3, 128, 24000, 85),
    ("cfg5 L1", "256x0e+256x1o+256x2e", 3, 256, 1600, 42),  # MACE large, 8 cells of ~200 atoms per GPU
]


def timed(fn):
    ts = []
    for _ in range(reps + 2):
        flush.zero_()
        torch.cuda._sleep(400000)  # the host runs ahead: the event pair brackets the kernel(s), not the Python dispatch
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts = sorted(ts[min(2, len(ts) - 1):])
    return ts[len(ts) // 2]


for name, node, lmax, C, N, deg in CASES:
    if only not in name:
        continue
    gen = torch.Generator().manual_seed(1)
    target = "+".join(f"{C}x{l}{'e' if l % 2 == 0 else 'o'}" for l in range(lmax + 1))
    plan = TPPlan(node, lmax, target)
    E = N * deg
    rcv = torch.arange(N).repeat_interleave(deg)
    snd = (rcv + torch.randint(1, min(N, 400), (E,), generator=gen)) % N
    order = torch.argsort(snd * N + rcv)  # sender-major like the neighbour list -> the receiver view needs perm
    ei = torch.stack([snd[order], rcv[order]]).to(dev)
    ep = ops.EdgePlan(ei, N)
    ei_g = ei[:, torch.argsort(ei[1], stable=True)]  # receiver-grouped order: perm = None, R is one sequential stream
    ep_g = ops.EdgePlan(ei_g, N, assume_grouped=True)
    h = torch.randn(N, C * plan.DIN, device=dev)
    Y = torch.randn(E, plan.NSH, device=dev)
    R = torch.randn(E, plan.NP * C, device=dev)
    w = torch.randn(N, C * plan.DOUT, device=dev)
    fwd_b = E * (4 * plan.NP * C + 4 * plan.NSH + 16) + N * 4 * C * (plan.DIN + plan.DOUT)
    bwd_b = E * (8 * plan.NP * C + 8 * plan.NSH + 16) + N * 4 * C * (2 * plan.DIN + plan.DOUT)
    for simple in modes:
        _lib.lib().pml_tp_set_simple(simple)
        tf = timed(lambda: ops.tp_scatter(h, Y, R, ep.gather, ep.rowptr, ep.perm, plan.handle))
        tb = timed(lambda: ops.tp_scatter_bwd(h, Y, R, w, ep.gather, ep.rowptr, ep.perm, plan.handle, True, True, True))
        tr = timed(lambda: ops.tp_scatter_bwd(h, Y, R, w, ep.gather, ep.rowptr, ep.perm, plan.handle, False, False, True))
        kind = "streaming" if simple & 1 else ("pipelined" if simple == 0 else f"pipe mode{simple}")
        print(f"{name} E={E} {kind:9s} fwd {tf * 1e3:8.1f} us {fwd_b / tf / 1e6:7.0f} GB/s ({100 * fwd_b / tf / 1e6 / peak:4.1f}%) | "
              f"bwd(h,Y,R) {tb * 1e3:8.1f} us {bwd_b / tb / 1e6:7.0f} GB/s ({100 * bwd_b / tb / 1e6 / peak:4.1f}%) | "
              f"bwd(R only) {tr * 1e3:8.1f} us (incl. allocs/memsets of the op wrapper)", flush=True)
    _lib.lib().pml_tp_set_simple(0)
    if len(modes) < 3:
        continue
    assert ep_g.perm is None
    tf = timed(lambda: ops.tp_scatter(h, Y, R, ep_g.gather, ep_g.rowptr, None, plan.handle))
    tb = timed(lambda: ops.tp_scatter_bwd(h, Y, R, w, ep_g.gather, ep_g.rowptr, None, plan.handle, True, True, True))
    print(f"{name} E={E} pipelined, receiver-grouped edges (no perm): fwd {tf * 1e3:8.1f} us ({100 * fwd_b / tf / 1e6 / peak:4.1f}%) | "
          f"bwd(h,Y,R) {tb * 1e3:8.1f} us ({100 * bwd_b / tb / 1e6 / peak:4.1f}%)", flush=True)
