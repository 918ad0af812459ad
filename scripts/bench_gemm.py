"""Micro-benchmark of the grouped GEMM kernels on the radial-MLP shapes (E rows).  CUDA-event timing, L2-sized inputs."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from physicsml_b200 import ops  # noqa: E402
from physicsml_b200.plans import LinearPlan  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 41570
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
ops.TC_MIN_ROWS[0] = 1


def timeit(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # us


for k, n in [(8, 64), (64, 64), (64, 512), (64, 1280)]:
    plan = LinearPlan(f"{k}x0e", f"{n}x0e")
    x = torch.randn(E, k, generator=g).to(dev)
    w = torch.randn(k * n, generator=g).to(dev)
    gy = torch.randn(E, n, generator=g).to(dev)
    for mode in ("ffma", "tc"):
        ops.GEMM_MODE[0] = mode
        t_f = timeit(lambda: ops.irrep_linear(x, w, plan.handle))
        t_x = timeit(lambda: ops.irrep_linear_bwd_x(gy, w, plan.handle))
        t_w = timeit(lambda: ops.irrep_linear_bwd_w(x, gy, plan.handle))
        fl = 2.0 * E * k * n
        by_f = 4.0 * (E * k + k * n + E * n)
        print(f"[{mode:4s}] E={E} {k:4d}->{n:4d}  fwd {t_f:8.1f} us ({fl / t_f / 1e6:7.2f} TFLOP/s, {by_f / t_f / 1e3:7.1f} GB/s)"
              f"   dx {t_x:8.1f} us   dW {t_w:8.1f} us")
