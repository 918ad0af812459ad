"""BASELINE configs[2]: NequIP (4 layers, 64 channels, lmax 2, parity) energy+forces and one force-matching training
step on the config-2 batch of synthetic organic molecules (32 molecules, ~1300 atoms, ~41.6 k edges): the shared fused
edge kernels (K4), the radial MLP (K3 / K3w), the element-indexed self connection (K5b) and the gate kernel.
usage: python scripts/bench_nequip.py [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from physicsml_b200 import workloads as wl  # noqa: E402
from physicsml_b200.mace import compute_forces_by_gradient  # noqa: E402
from physicsml_b200.nequip import PooledNequipModule  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = torch.device("cuda:0")
data, targets, kw, _ = wl.make_batch("mace_medium_spice_train", 0)
Z = data["node_attrs"].shape[1]
torch.manual_seed(0)
model = PooledNequipModule(mlp_irreps="16x0e", cut_off=kw["cut_off"], num_layers=4, max_ell=2, parity=True, num_features=64,
                           num_bessel=8, bessel_basis_trainable=False, num_polynomial_cutoff=6, self_connection=True,
                           resnet=True, num_node_feats=Z, num_edge_feats=0,
                           avg_num_neighbours=wl.mean_degree("mace_medium_spice_train")).to(dev)
opt = wl.make_optimizer(model)
d = {k: (v.to(dev) if torch.is_tensor(v) and v.dim() > 0 else v) for k, v in data.items()}
t = {k: v.to(dev) for k, v in targets.items()}
n_atoms, n_edges = d["coordinates"].shape[0], d["edge_index"].shape[1]


def infer():
    dd = dict(d)
    dd["coordinates"] = d["coordinates"].detach().requires_grad_(True)
    e = model(dd)["y_graph_scalars"]
    return compute_forces_by_gradient(e, dd["coordinates"], training=False)


def train():
    opt.zero_grad(set_to_none=True)
    loss = wl.training_loss(model, d, t)
    loss.backward()
    opt.step()


def timeit(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


ti, tt = timeit(infer), timeit(train)
print(f"NequIP 4 layers x 64 features, lmax 2: {n_atoms} atoms, {n_edges} edges, eager (no CUDA graph; host-dispatch-bound)")
print(f"  energy+forces  {ti:8.2f} ms -> {n_atoms / ti * 1e3:10.0f} atoms/s")
print(f"  training step  {tt:8.2f} ms -> {n_atoms / tt * 1e3:10.0f} atoms/s")
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        train()
    torch.cuda.synchronize()
tot = sum(e.self_device_time_total for e in prof.key_averages()) / 3e3
print(f"  device time of one training step (sum of kernels, CUPTI): {tot:.2f} ms")
print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=14, max_name_column_width=60))
