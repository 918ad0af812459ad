"""Where does a training step go?  Host enqueue time vs device time, and the per-kernel device-time table
(torch.profiler / CUPTI; no ncu needed).  Usage: python scripts/profile_step.py [workload]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from physicsml_b200 import workloads as wl  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "mace_medium_spice_train"
dev = torch.device("cuda:0")
data, targets, kw, mlp = wl.make_batch(name, 0)
model = wl.build_model(name).to(dev)
opt = wl.make_optimizer(model)
d = {k: (v.to(dev) if torch.is_tensor(v) and v.dim() > 0 else v) for k, v in data.items()}
t = {k: v.to(dev) for k, v in targets.items()}


def step():
    opt.zero_grad(set_to_none=True)
    loss = wl.training_loss(model, d, t)
    loss.backward()
    opt.step()


for _ in range(5):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3 * (t1 - t0) / 10:.2f} ms/step ; wall incl. drain {1e3 * (t2 - t0) / 10:.2f} ms/step")
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=70, max_name_column_width=70))

if os.environ.get("PML_PROFILE_SHAPES") == "1":  # which eager ATen ops (and tensor shapes) are still on the device
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
        step()
        torch.cuda.synchronize()
    rows = [e for e in prof.key_averages(group_by_input_shape=True) if e.key.startswith("aten::") and e.self_device_time_total > 0]
    rows.sort(key=lambda e: -e.self_device_time_total)
    print("\nATen ops with device time, one step (op, input shapes, calls, device us):")
    for e in rows[:40]:
        print(f"{e.key:28s} {str(e.input_shapes)[:90]:90s} {e.count:4d} {e.self_device_time_total:9.1f}")
    print(f"total ATen device time {sum(e.self_device_time_total for e in rows):.1f} us")
    big = [e for e in prof.events() if e.name.startswith("aten::") and e.self_device_time_total > 20 and e.stack]
    seen = set()
    for e in big:
        frames = [s for s in e.stack if "physicsml_b200" in s or "workloads" in s][:3]
        key = (e.name, tuple(frames))
        if key in seen:
            continue
        seen.add(key)
        print(e.name, [list(s) for s in (e.input_shapes or [])][:3], f"{e.self_device_time_total:.0f}us", frames)
