"""Microbenchmark of the K6 symmetric-contraction kernels (fwd / bwd / dbl) at config-2 size.

usage: python scripts/bench_sym.py [reps] [N] [C]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from physicsml_b200 import mace, ops  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1304
Cc = int(sys.argv[3]) if len(sys.argv) > 3 else 128
torch.manual_seed(0)
dev = torch.device("cuda")
Z = 10
ir_in = " + ".join(f"{Cc}x{l}{'e' if l % 2 == 0 else 'o'}" for l in range(4))
for ir_out in (f"{Cc}x0e + {Cc}x1o", f"{Cc}x0e"):
    sc = mace.SymmetricContraction(ir_in, ir_out, 3, Z).to(dev)
    pl = sc.plan
    S = 16
    x = torch.randn(N, Cc, S, device=dev)
    v = torch.randn_like(x)
    z = torch.randint(0, 4, (N,), device=dev)
    attrs = torch.nn.functional.one_hot(z, Z).float()
    ws = [w.detach() for w in pl.reduce_weights([w for c in sc.contractions for w in c.by_nu()])]
    g = torch.randn(N, pl.C * pl.DOUT, device=dev)

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

    t_f = timeit(lambda: ops.sym_contract(x, attrs, ws, pl.handle))
    t_bx = timeit(lambda: ops.sym_contract_bwd(x, attrs, ws, g, pl.handle, True, False))
    t_bw = timeit(lambda: ops.sym_contract_bwd(x, attrs, ws, g, pl.handle, True, True))
    t_d = timeit(lambda: ops.sym_contract_dbl(x, v, attrs, ws, g, pl.handle, True, True, False))
    t_dw = timeit(lambda: ops.sym_contract_dbl(x, v, attrs, ws, g, pl.handle, True, True, True))
    print(f"{ir_out:>20s} N={N} C={Cc} K={[int(w.shape[1]) for w in ws]}: fwd {t_f:7.1f} us | bwd(x) {t_bx:7.1f} | "
          f"bwd(x,W) {t_bw:7.1f} | dbl(g,x) {t_d:7.1f} | dbl(g,x,W) {t_dw:7.1f}  (incl. op-wrapper allocs)")
