"""Fixed cost vs per-k-block cost of the tensor-core GEMM on node-sized problems (1 304 rows): y = x W for K in
(16 .. 512) and N in (32, 128), 50 launches captured in one CUDA graph and replayed (no host launch gaps).
Usage: python scripts/bench_small_gemm.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from physicsml_b200 import ops  # noqa: E402
from physicsml_b200.plans import LinearPlan  # noqa: E402

dev = torch.device("cuda:0")
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1304
REP = 50
for N in (32, 128):
    for K in (16, 32, 64, 128, 256, 512):
        pl = LinearPlan(f"{K}x0e", f"{N}x0e")
        x = torch.randn(rows, K, device=dev)
        w = torch.randn(K * N, device=dev)
        ys = [ops.irrep_linear(x, w, pl.handle) for _ in range(3)]
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            y = x
            for _ in range(REP):
                y = ops.irrep_linear(x, w, pl.handle)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        print(f"rows {rows} K {K:4d} N {N:4d}: {1e3 * e0.elapsed_time(e1) / (10 * REP):7.2f} us per launch (independent launches, in-graph)")
