"""Wide radial-layer GEMMs (csrc/gemm_wide.cu) vs the general grouped tensor-core kernel: accuracy against fp64 and
CUDA-event time.  usage: python scripts/bench_wide.py [reps] [rows]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from physicsml_b200 import ops, plans  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
E = int(sys.argv[2]) if len(sys.argv) > 2 else 41570
dev = torch.device("cuda")
torch.manual_seed(0)


def timeit(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    torch.cuda._sleep(400000)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


for K, N in ((64, 1280), (64, 512), (64, 64), (8, 64)):
    pl = plans.LinearPlan(f"{K}x0e", f"{N}x0e")
    x = torch.randn(E, K, device=dev)
    w = torch.randn(K * N, device=dev)
    g = torch.randn(E, N, device=dev)
    ref_y = (x.double() @ w.double().view(K, N)) / K ** 0.5
    ref_dx = (g.double() @ w.double().view(K, N).t()) / K ** 0.5
    ref_dw = (x.double().t() @ g.double()).reshape(-1) / K ** 0.5
    res = {}
    for mode, rows in (("general", 1 << 30), ("wide", 4096)):
        ops.WIDE_MIN_ROWS[0] = rows
        y = ops.irrep_linear(x, w, pl.handle)
        dx = ops.irrep_linear_bwd_x(g, w, pl.handle)
        dw = ops.irrep_linear_bwd_w(x, g, pl.handle)
        err = [((a.double() - b).abs().max() / b.abs().max()).item() for a, b in ((y, ref_y), (dx, ref_dx), (dw, ref_dw))]
        t = [timeit(lambda: ops.irrep_linear(x, w, pl.handle)), timeit(lambda: ops.irrep_linear_bwd_x(g, w, pl.handle)),
             timeit(lambda: ops.irrep_linear_bwd_w(x, g, pl.handle))]
        hbm = [4 * E * (K + N) / 1e3, 4 * E * (K + N) / 1e3, 4 * E * (K + N) / 1e3]  # KB
        print(f"{E} x {K} -> {N} {mode:8s}: fwd {t[0]:7.1f} us ({hbm[0] / t[0]:6.0f} GB/s) err {err[0]:.1e} | dx {t[1]:7.1f} us "
              f"({hbm[1] / t[1]:6.0f} GB/s) err {err[1]:.1e} | dW {t[2]:7.1f} us ({hbm[2] / t[2]:6.0f} GB/s) err {err[2]:.1e}")
