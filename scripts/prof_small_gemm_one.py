"""one tensor-core GEMM launch of the small-grid variant (148 row tiles x K = 128 x N = 128) for an ncu source-level capture:
ncu --set full --warp-sampling-interval 0 --import-source on -k regex:gemm_grouped_tc -s 3 -c 1 python scripts/prof_small_gemm_one.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from physicsml_b200 import ops  # noqa: E402
from physicsml_b200.plans import LinearPlan  # noqa: E402

dev = torch.device("cuda:0")
pl = LinearPlan("128x0e", "128x0e")
x = torch.randn(18944, 128, device=dev)  # 148 row tiles: one CTA per SM
w = torch.randn(128 * 128, device=dev)
for _ in range(6):
    y = ops.irrep_linear(x, w, pl.handle)
torch.cuda.synchronize()
