"""one forward / dx / dW launch of the tensor-core GEMM at a radial-MLP shape (for ncu)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from physicsml_b200 import ops  # noqa: E402
from physicsml_b200.plans import LinearPlan  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 41570
k, n = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (64, 1280)
dev = torch.device("cuda:0")
plan = LinearPlan(f"{k}x0e", f"{n}x0e")
x = torch.randn(E, k, device=dev)
w = torch.randn(k * n, device=dev)
gy = torch.randn(E, n, device=dev)
for _ in range(3):
    ops.irrep_linear(x, w, plan.handle)
    ops.irrep_linear_bwd_x(gy, w, plan.handle)
    ops.irrep_linear_bwd_w(x, gy, plan.handle)
torch.cuda.synchronize()
