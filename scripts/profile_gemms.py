"""Per-shape device time of every grouped-GEMM launch (K3) in one training step: wraps ops._gemm with CUDA events.
Usage: python scripts/profile_gemms.py [workload]"""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from physicsml_b200 import ops  # noqa: E402
from physicsml_b200 import workloads as wl  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "mace_medium_spice_train"
dev = torch.device("cuda:0")
data, targets, kw, mlp = wl.make_batch(name, 0)
model = wl.build_model(name).to(dev)
opt = wl.make_optimizer(model)
d = {k: (v.to(dev) if torch.is_tensor(v) and v.dim() > 0 else v) for k, v in data.items()}
t = {k: v.to(dev) for k, v in targets.items()}
records = []
orig = ops._gemm
recording = [False]


def wrapped(gname, a, b, c, table, nprob, m_runtime, max_m, max_n, max_batch, splitk, max_k):
    if not recording[0]:
        return orig(gname, a, b, c, table, nprob, m_runtime, max_m, max_n, max_batch, splitk, max_k)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    orig(gname, a, b, c, table, nprob, m_runtime, max_m, max_n, max_batch, splitk, max_k)
    e1.record()
    caller = sys._getframe(1).f_code.co_name
    records.append(((caller, gname.replace("pml_gemm_grouped", "g"), nprob, m_runtime, max_m, max_n, max_batch, splitk, max_k), e0, e1))


ops._gemm = wrapped


def step():
    opt.zero_grad(set_to_none=True)
    loss = wl.training_loss(model, d, t)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
recording[0] = True
NS = 3
for _ in range(NS):
    step()
torch.cuda.synchronize()
agg = collections.OrderedDict()
for key, e0, e1 in records:
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += e0.elapsed_time(e1) * 1e3
tot = sum(v[1] for v in agg.values()) / NS
print(f"total GEMM time {tot:.0f} us/step over {sum(v[0] for v in agg.values()) // NS} launches/step (event-bracketed, includes launch gaps)")
print("caller, kernel, nprob, M_runtime, max_M, max_N, batch, splitk, max_k : launches/step, us each, us/step")
for key, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{str(key):100s} {n / NS:5.1f} {us / n:8.1f} {us / NS:8.0f}")
