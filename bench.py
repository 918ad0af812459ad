#!/usr/bin/env python
"""Benchmark of the MACE hot path (BASELINE.json metric: MACE energy+forces atoms/s; fused-edge-kernel % HBM roofline).

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU cores (oracle port)

Workload at every N: configs[1] of BASELINE.json — MACE medium (128x0e+128x1o, lmax 3, correlation 3) TRAINING step on
32 synthetic SPICE-like molecules (30-50 atoms) per GPU: forward, forces with create_graph, MSE(E)+MSE(F), backward,
gradient all-reduce (NCCL, N > 1), Adam-amsgrad.  Weak scaling: per-GPU work is fixed.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOAD = "mace_medium_spice_train"
METRIC = "mace_energy_forces_atoms_per_sec"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int) -> None:
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------------
def cpu_reference_step_rate(n_mol: int, steps: int, warmup: int):
    """Reference algorithm (oracle restatement of the reference's PyTorch/e3nn path) on the host cores: same model,
    same training step, on a bounded sample of the workload (n_mol molecules).  Returns (atoms/s, cores, sample str)."""
    import torch

    from oracle.ref_model import RefPooledMACE
    from physicsml_b200 import workloads as wl

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    data, targets, kw, mlp = wl.make_batch(WORKLOAD, 0, n_mol=n_mol)
    kw["avg_num_neighbours"] = wl.mean_degree(WORKLOAD)
    torch.manual_seed(0)
    model = RefPooledMACE(mlp_irreps=mlp, **kw)
    opt = wl.make_optimizer(model)
    n_atoms = data["coordinates"].shape[0]
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        loss = wl.training_loss(model, data, targets)
        loss.backward()
        opt.step()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    per = sum(times) / len(times)
    sample = f"{n_mol} of 32 molecules ({n_atoms} atoms, {data['edge_index'].shape[1]} edges), {steps} timed training steps, fp32, {cores} threads"
    return n_atoms / per, cores, sample, per * 1e3


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_mol = args.ref_molecules
    rate, cores, sample, ms = cpu_reference_step_rate(n_mol, max(1, args.steps), max(1, args.warmup))
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "atoms/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": rate, "unit": "atoms/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "atoms/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------------------
def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-molecules", type=int, default=4, help="bounded sample size of the CPU reference legs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs only)")
    ap.add_argument("--eager", action="store_true", help="do not capture the step in a CUDA graph")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    from physicsml_b200 import ops, plans
    from physicsml_b200 import workloads as wl
    from physicsml_b200.mace import PooledMACEModule

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: physicsml_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cuda.matmul.allow_tf32 = False

    data, targets, kw, mlp = wl.make_batch(WORKLOAD, rank)
    kw["avg_num_neighbours"] = wl.mean_degree(WORKLOAD)
    torch.manual_seed(0)
    model = PooledMACEModule(mlp_irreps=mlp, **kw).to(dev)
    use_graph = not args.eager
    opt = wl.make_optimizer(model, capturable=use_graph)
    params = [p for p in model.parameters() if p.requires_grad and p.numel() > 0]
    n_atoms = data["coordinates"].shape[0]
    n_edges = data["edge_index"].shape[1]

    # host (pinned) copies for the end-to-end leg, device-resident copies for the kernel-path leg
    host = {k: (v.pin_memory() if torch.is_tensor(v) and v.dim() > 0 else v) for k, v in data.items()}
    host_t = {k: v.pin_memory() for k, v in targets.items()}
    d_dev = {k: (v.to(dev) if torch.is_tensor(v) and v.dim() > 0 else v) for k, v in data.items()}
    t_dev = {k: v.to(dev) for k, v in targets.items()}
    h2d = sum(v.numel() * v.element_size() for v in host.values() if torch.is_tensor(v) and v.dim() > 0)
    h2d += sum(v.numel() * v.element_size() for v in host_t.values())
    loss_host = torch.zeros(1).pin_memory()

    def allreduce_grads():
        if world == 1:
            return
        flat = torch.cat([p.grad.reshape(-1) for p in params])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(world)
        off = 0
        for p in params:
            n = p.numel()
            p.grad.copy_(flat[off : off + n].view_as(p))
            off += n

    def step(d, t):
        opt.zero_grad(set_to_none=True)
        loss = wl.training_loss(model, d, t)
        loss.backward()
        allreduce_grads()
        opt.step()
        return loss

    graphed = None
    if use_graph:
        from physicsml_b200.graphs import GraphedStep

        try:
            graphed = GraphedStep(model, opt, d_dev, t_dev, wl.training_loss, after_backward=allreduce_grads, warmup=3)
        except Exception as exc:  # capture is an optimisation: report and continue eagerly
            print(f"[bench] CUDA-graph capture failed ({type(exc).__name__}: {exc}); running eagerly", file=sys.stderr)
            graphed = None
            opt = wl.make_optimizer(model, capturable=False)

    def step_resident():
        if graphed is not None:
            return graphed()
        return step(d_dev, t_dev)

    def step_e2e():
        if graphed is not None:
            graphed.load(host, host_t)  # pinned host -> static device buffers
            loss = graphed()
        else:
            d = {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) and v.dim() > 0 else v) for k, v in host.items()}
            t = {k: v.to(dev, non_blocking=True) for k, v in host_t.items()}
            loss = step(d, t)
        loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the caller reads the loss every step
        return float(loss_host[0])

    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, with_kernel_timing=False):
        for _ in range(warmup):
            fn()
        barrier()
        if with_kernel_timing:
            ops.enable_kernel_timing(["pml_tp_fwd", "pml_tp_bwd"])
        n0 = ops.LAUNCHES[0]
        evs = []
        for _ in range(steps):
            flush.zero_()  # L2 flush (256 MiB > 126 MB L2), outside the per-step event pair
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            evs.append((e0, e1))
        barrier()
        launches = ops.LAUNCHES[0] - n0
        ms = sum(a.elapsed_time(b) for a, b in evs) / steps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        return ms, launches

    sampler = ClockSampler(local) if rank == 0 else None
    ms_dev, launches = timed(step_resident, args.steps, args.warmup, with_kernel_timing=(graphed is None))
    clocks = sampler.stop() if sampler is not None else None
    if graphed is not None:
        # a replayed graph cannot carry per-kernel events: time the fused edge kernels in an eager pass of the SAME step
        # (same inputs, same kernels), and count the kernels of one step the same way
        opt_e = wl.make_optimizer(model, capturable=False)
        opt_saved, opt = opt, opt_e
        _, launches = timed(lambda: step(d_dev, t_dev), args.steps, 3, with_kernel_timing=True)
        opt = opt_saved
    # per-kernel times of the fused edge kernels, recorded live inside the timed region
    ktimes: dict = {}
    for name, recs in ops.KERNEL_TIMING.items():
        for spec, a, b in recs:
            ktimes.setdefault((name, spec), []).append(a.elapsed_time(b))
    ops.enable_kernel_timing([])
    ms_e2e = float("nan")
    if not args.no_e2e:
        ms_e2e, _ = timed(step_e2e, args.steps, args.warmup)

    total_atoms = n_atoms
    if world > 1:
        t = torch.tensor([float(n_atoms)], device=dev)
        dist.all_reduce(t)
        total_atoms = float(t[0])

    def shutdown():
        """Leave the process group without hanging: the captured step holds NCCL work, so drop the graph first; a
        watchdog forces the exit if the communicator teardown still blocks (the JSON line is already flushed)."""
        nonlocal graphed
        if world == 1:
            return
        import gc
        import threading

        sys.stdout.flush()
        threading.Timer(30.0, lambda: os._exit(0)).start()
        graphed = None
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
        os._exit(0)

    if rank != 0:
        shutdown()
        return

    # roofline of the dominant fused edge kernel (forward, the layer with the most traffic)
    peak, which = _peaks()
    roof = None
    best = None
    for (name, spec), ts in ktimes.items():
        if name != "pml_tp_fwd":
            continue
        pl = next(p for p in plans._PLANS if getattr(p, "spec_id", None) == spec and hasattr(p, "NP") and hasattr(p, "DIN"))
        # bytes_fwd = E*(4*P*C + 4*S + 16) + N*4*(d_in + d_tp)
        byts = n_edges * (4 * pl.NP * pl.C + 4 * pl.NSH + 16) + n_atoms * 4 * (pl.C * pl.DIN + pl.C * pl.DOUT)
        avg_ms = sum(ts) / len(ts)
        if best is None or byts > best[0]:
            best = (byts, avg_ms, len(ts), pl.spec.key)
    if best is not None:
        byts, avg_ms, n, key = best
        ach = byts / (avg_ms * 1e-3) / 1e9
        # dram__bytes_read.sum + dram__bytes_write.sum of this kernel on this workload from the committed ncu --set full
        # capture (profiles/README.md says which command produced it); null when no capture of this signature exists
        traffic, tsrc = None, None
        tpath = os.path.join(ROOT, "profiles", "k4_dram_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                rec = json.load(f).get(f"{WORKLOAD}:tp_fwd:{key}")
            if rec:
                traffic, tsrc = rec["dram_bytes_per_launch"], rec["source"]
        roof = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                "traffic_source": tsrc, "kernel": f"tp_fwd_pipe_kernel<{key}>", "algorithmic_bytes_per_launch": byts,
                "avg_launch_ms": avg_ms, "launches_timed": n, "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({which})"}
    kernel_share = {f"{n}[{s}]": {"launches": len(ts), "total_ms_per_step": sum(ts) / args.steps} for (n, s), ts in ktimes.items()}

    cpu = None
    if not args.no_cpu_baseline and world == 1:  # reported on rank 0 at N = 1 only
        rate, cores, sample, _ = cpu_reference_step_rate(args.ref_molecules, 2, 1)
        cpu = {"value": rate, "unit": "atoms/s", "cores": cores, "kind": "port", "sample": sample}

    line = {
        "metric": METRIC, "value": total_atoms / (ms_dev * 1e-3), "unit": "atoms/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "step": "training: fwd + forces(create_graph) + MSE(E)+MSE(F) + bwd + allreduce + Adam(amsgrad, fused kernel)",
                   "molecules_per_gpu": 32, "atoms_per_gpu": n_atoms, "edges_per_gpu": n_edges, "parallelism": f"dp{world}",
                   "l2": "flushed between steps (256 MiB memset outside the per-step event pair)"},
        "e2e": {"value": total_atoms / (ms_e2e * 1e-3), "unit": "atoms/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e},
        "gpu_launches": launches,
        "cuda_graph": graphed is not None,
        "clocks": clocks,
        "roofline": roof,
        "cpu_baseline": cpu,
        "edge_kernels": kernel_share,
    }
    print(json.dumps(line), flush=True)
    shutdown()


if __name__ == "__main__":
    main()
