"""Where does an MD step of the periodic water box go?  Per step: host wall clock, device time of the evaluation (CUDA
events), cudaMalloc calls and allocator retries (a neighbour list of a different length every step changes every
edge-sized allocation).  Usage: python scripts/md_probe.py [steps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from physicsml_b200 import workloads as wl  # noqa: E402
from physicsml_b200.md import MDCalculator  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
name = "water_box_infer"
dev = torch.device("cuda:0")
data, _, kw, _ = wl.make_batch(name, 0)
model = wl.build_model(name).to(dev).eval()
cfg = wl.CONFIGS[name]
md = MDCalculator(model, cfg["model"]["cut_off"], data["node_attrs"], cell=data["cell"])
n = int(data["coordinates"].shape[0])
gen = torch.Generator().manual_seed(5)
pos = data["coordinates"].clone()
host_pos = torch.empty(n, 3).pin_memory()
for i in range(steps + 3):
    pos = pos + 0.017 * torch.randn(n, 3, generator=gen)
    host_pos.copy_(pos)
    torch.cuda.synchronize()
    s0 = torch.cuda.memory_stats()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    e, f = md.energy_forces(host_pos)
    ev1.record()
    t1 = time.perf_counter()
    fh = f.cpu()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    s1 = torch.cuda.memory_stats()
    print(f"step {i}: wall {1e3 * (t2 - t0):7.2f} ms  host enqueue {1e3 * (t1 - t0):7.2f} ms  device {ev0.elapsed_time(ev1):7.2f} ms  "
          f"edges {md.n_edges}  cudaMalloc +{s1['num_device_alloc'] - s0['num_device_alloc']} cudaFree +{s1['num_device_free'] - s0['num_device_free']} "
          f"retries +{s1['num_alloc_retries'] - s0['num_alloc_retries']}  reserved {s1['reserved_bytes.all.current'] / 2**30:.1f} GiB")
