"""Which Python lines launch the small torch kernels (fills, adds, copies) of one training step?
TorchDispatchMode over one eager step; every aten op that launches a kernel is attributed to the innermost frame inside
this repository.  Usage: python scripts/count_small_ops.py [workload]"""
import collections
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.utils._python_dispatch import TorchDispatchMode  # noqa: E402

from physicsml_b200 import workloads as wl  # noqa: E402
from physicsml_b200.mace import PooledMACEModule  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "mace_medium_spice_train"
dev = torch.device("cuda:0")
data, targets, kw, mlp = wl.make_batch(name, 0)
kw["avg_num_neighbours"] = wl.mean_degree(name)
torch.manual_seed(0)
model = PooledMACEModule(mlp_irreps=mlp, **kw).to(dev)
opt = wl.make_optimizer(model)
d = {k: (v.to(dev) if torch.is_tensor(v) and v.dim() > 0 else v) for k, v in data.items()}
t = {k: v.to(dev) for k, v in targets.items()}
SKIP = ("aten.view", "aten._unsafe_view", "aten.reshape", "aten.t.", "aten.transpose", "aten.permute", "aten.select",
        "aten.slice", "aten.unsqueeze", "aten.squeeze", "aten.expand", "aten.detach", "aten.alias", "aten.as_strided",
        "aten.empty", "aten.new_empty", "aten.unbind", "aten.split", "aten.lift_fresh", "aten.is_", "aten.sym_", "aten.stride",
        "aten.size", "aten._local_scalar", "aten.narrow", "aten.empty_like", "aten.new_empty_strided", "aten.unflatten")
counts = collections.Counter()


class Mode(TorchDispatchMode):
    def __torch_dispatch__(self, func, types, args=(), kwargs=None):
        n = str(func)
        if not n.startswith("physicsml_b200") and not any(n.startswith(s) for s in SKIP):
            where = "?"
            for fr in reversed(traceback.extract_stack()[:-1]):
                if fr.filename.startswith(ROOT) and "count_small_ops" not in fr.filename:
                    where = f"{os.path.relpath(fr.filename, ROOT)}:{fr.lineno} {fr.name}"
                    break
            counts[(n, where)] += 1
        return func(*args, **(kwargs or {}))


def step():
    opt.zero_grad(set_to_none=True)
    loss = wl.training_loss(model, d, t)
    loss.backward()
    opt.step()


for _ in range(2):
    step()
with Mode():
    step()
torch.cuda.synchronize()
by_op = collections.Counter()
for (n, w), c in counts.items():
    by_op[n] += c
print("per op:", dict(by_op.most_common(20)))
for (n, w), c in counts.most_common(70):
    print(f"{c:4d}  {n:32s} {w}")
