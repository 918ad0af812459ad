#!/usr/bin/env python
"""Benchmark of the MACE / NequIP message-passing hot path (BASELINE.json metric: MACE energy+forces atoms/s at
1/2/4/8 B200; fused-edge-kernel % HBM roofline).

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own code on the host CPU cores

Headline workload at every N: configs[1] of BASELINE.json — MACE medium (128x0e+128x1o, lmax 3, correlation 3) TRAINING
step on 32 synthetic SPICE-like molecules (30-50 atoms) per GPU: forward, forces with create_graph, MSE(E)+MSE(F),
backward, gradient all-reduce (NCCL, N > 1), Adam-amsgrad.  Weak scaling: per-GPU work is fixed.  The other four
BASELINE configs are measured in the same run and printed under ``other_configs`` (training configs at every N,
inference configs at N = 1).  Prints ONE JSON line on rank 0.
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
OTHERS = ["mace_small_qm9", "nequip_organic", "water_box_infer", "mace_large_bulk_train"]
METRIC = "mace_energy_forces_atoms_per_sec"
PARITY_MOLECULES = 4  # sub-batch the GPU arm is checked on against the CPU oracle inside the benchmark run


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def shared_config(world: int = 1) -> dict:
    """the `config` object both arms print (same workload, same step; the driver compares them)"""
    from physicsml_b200 import workloads as wl

    data, _, _, _ = wl.make_batch(WORKLOAD, 0)
    return {"workload": WORKLOAD,
            "step": "training: fwd + forces(create_graph) + MSE(E)+MSE(F) + bwd + allreduce + Adam(amsgrad)",
            "molecules_per_gpu": wl.CONFIGS[WORKLOAD]["n_mol"], "atoms_per_gpu": int(data["coordinates"].shape[0]),
            "edges_per_gpu": int(data["edge_index"].shape[1]),
            "per_rank_batches": "same molecule sizes on every rank (fixed work per GPU), different coordinates and species",
            "l2": "GPU arm: L2 flushed between timed steps (256 MiB memset outside the per-step CUDA-event pair)"}


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
        # median over the samples taken while the GPU was clocked up (idle samples between legs are not "under load")
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------------
# CPU legs: the reference's own code (oracle/_ref, unmodified sources over the e3nn shims) or, if that was not staged,
# the oracle port
# ----------------------------------------------------------------------------------------------------------------------
def _cpu_model(name: str):
    """-> (model, kind): the reference's own modules when oracle/_ref (or /root/reference) is present, else the port"""
    import torch

    from physicsml_b200 import workloads as wl

    cfg = wl.CONFIGS[name]
    try:
        from oracle import reference_models as rm

        if rm.available():
            kw = dict(cfg["model"])
            kw["avg_num_neighbours"] = wl.mean_degree(name)
            torch.manual_seed(0)
            cls = rm.ReferencePooledMACE if cfg["family"] == "mace" else rm.ReferencePooledNequip
            return cls(cfg["mlp"], **kw), "reference"
    except Exception as exc:  # fall back to the port, loudly
        print(f"[bench] reference sources unavailable ({type(exc).__name__}: {exc}); timing the oracle port", file=sys.stderr)
    return wl.build_model(name, oracle=True), "port"


def cpu_reference_step_rate(n_mol, steps: int, warmup: int, budget_s: float | None = None):
    """The reference's training step on the host cores, same model, same step, on `n_mol` molecules of the workload
    (None = the full batch).  -> (atoms/s, cores, sample str, ms per step, kind)"""
    import torch

    from physicsml_b200 import workloads as wl

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    data, targets, _, _ = wl.make_batch(WORKLOAD, 0, n_mol=n_mol)
    model, kind = _cpu_model(WORKLOAD)
    opt = wl.make_optimizer(model)
    n_atoms = data["coordinates"].shape[0]
    times = []
    t_start = time.perf_counter()
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        loss = wl.training_loss(model, data, targets)
        loss.backward()
        opt.step()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if budget_s is not None and times and time.perf_counter() - t_start > budget_s:
            break
    per = sum(times) / len(times)
    total = wl.CONFIGS[WORKLOAD]["n_mol"]
    sample = (f"{n_mol or total} of {total} molecules ({n_atoms} atoms, {data['edge_index'].shape[1]} edges), {len(times)} timed "
              f"training steps, fp32, {cores} threads")
    return n_atoms / per, cores, sample, per * 1e3, kind


def gpu_eager_step_rate(dev, steps: int = 5, warmup: int = 2) -> dict:
    """the headline training step with the reference's own modules (or the oracle port) moved to the GPU: stock PyTorch
    kernels, stock autograd (create_graph for the forces), torch.optim.Adam.  None of this repo's kernels on the path."""
    import torch

    from physicsml_b200 import workloads as wl

    data, targets, _, _ = wl.make_batch(WORKLOAD, 0)
    model, kind = _cpu_model(WORKLOAD)
    model = model.to(dev)
    d = {k: (v.to(dev) if torch.is_tensor(v) and v.dim() > 0 else v) for k, v in data.items()}
    t = {k: v.to(dev) for k, v in targets.items()}
    opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad and p.numel() > 0], lr=1e-2, amsgrad=True)

    def step():
        opt.zero_grad(set_to_none=True)
        dd = dict(d)
        pos = dd["coordinates"].detach().requires_grad_(True)
        dd["coordinates"] = pos
        e = model(dd)["y_graph_scalars"]
        (g,) = torch.autograd.grad([e], [pos], grad_outputs=[torch.ones_like(e)], retain_graph=True, create_graph=True)
        loss = torch.nn.functional.mse_loss(e, t["energy"]) + torch.nn.functional.mse_loss(-g, t["forces"])
        loss.backward()
        opt.step()

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    n = int(d["coordinates"].shape[0])
    return {"value": n / (ms * 1e-3), "unit": "atoms/s", "ms_per_step": ms, "kind": kind + " modules, eager PyTorch on the same GPU",
            "steps": steps}


def _config_molecules() -> int:
    from physicsml_b200 import workloads as wl

    return int(wl.CONFIGS[WORKLOAD]["n_mol"])


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # N = 1 (the line the headline ratio is computed from): the FULL 32-molecule batch, ~19 s per step on 16 cores.  In the
    # scaling run (N > 1) the reference arm is still one CPU process on rank 0: there it runs a bounded 4-molecule sample
    # of the same workload so that four more runs of K + W steps stay within minutes (cpu_baseline.sample says which)
    n_mol = args.ref_molecules if args.ref_molecules > 0 else (None if args.gpus <= 1 else PARITY_MOLECULES)
    if n_mol is None:
        # the whole K + W run has to end within a few minutes: one full-batch step of the reference costs ~19 s on 16 host
        # cores (0.6 s per molecule, measured), so the full batch is run when K + W <= 12 steps fit the budget and the
        # largest prefix of the batch that does otherwise (cpu_baseline.sample says which; atoms/s barely depends on it)
        total = _config_molecules()
        fit = int(args.ref_budget_s / (max(1, args.steps) + max(1, args.warmup)) / 0.6)
        n_mol = None if fit >= total else max(PARITY_MOLECULES, fit)
    rate, cores, sample, ms, kind = cpu_reference_step_rate(n_mol, max(1, args.steps), max(1, args.warmup))
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "atoms/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": shared_config(),
        "cpu_baseline": {"value": rate, "unit": "atoms/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": rate, "unit": "atoms/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------------
def edge_kernel_bytes(pl, n_nodes: int, n_edges: int, mask=None) -> int:
    """ALGORITHMIC bytes of one fused edge-kernel launch: every tensor once.
    forward: E*(4*P*C + 4*S + 16) + N*4*C*(DIN + DOUT).  cotangent with outputs mask = (dh, dY, dR): reads Y, the
    indices, h and g rows, R unless only dR is wanted; writes the requested cotangents."""
    PC, S = pl.NP * pl.C, pl.NSH
    if mask is None:
        return n_edges * (4 * PC + 4 * S + 16) + n_nodes * 4 * pl.C * (pl.DIN + pl.DOUT)
    need_h, need_y, need_r = mask
    b = n_edges * (4 * S + 16) + n_nodes * 4 * pl.C * (pl.DIN + pl.DOUT)
    if need_h or need_y:
        b += n_edges * 4 * PC
    if need_r:
        b += n_edges * 4 * PC
    if need_y:
        b += n_edges * 4 * S
    if need_h:
        b += n_nodes * 4 * pl.C * pl.DIN
    return b


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-molecules", type=int, default=0,
                    help="molecules of the batch the --impl reference arm runs (0 = the full 32-molecule batch at --gpus 1, a "
                         "4-molecule sample in the scaling runs)")
    ap.add_argument("--ref-budget-s", type=float, default=240.0,
                    help="wall-clock budget of the --impl reference run at --gpus 1: the sample of the batch is sized to it")
    ap.add_argument("--cpu-baseline-molecules", type=int, default=PARITY_MOLECULES,
                    help="bounded sample of the cpu_baseline leg inside the GPU arm's run")
    ap.add_argument("--configs", default="all", help="'all', 'headline' or a comma-separated list of other_configs to run")
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

    from physicsml_b200 import ops, parallel, plans
    from physicsml_b200 import workloads as wl
    from physicsml_b200.graphs import GraphedEval, GraphedStep

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
    peak, which = _peaks()
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    keep_alive = []  # captured graphs holding NCCL work are dropped explicitly at shutdown

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, kernel_timing=False):
        for _ in range(warmup):
            fn()
        barrier()
        if kernel_timing:
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
        launches = (ops.LAUNCHES[0] - n0) // max(1, steps)
        ms = sum(a.elapsed_time(b) for a, b in evs) / steps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        return ms, launches

    def edge_kernel_table(n_atoms, n_edges, steps):
        """per (kernel, signature, requested cotangents): launches per step, average launch time (CUDA events around each
        launch of the eager pass of the SAME step), algorithmic bytes, achieved GB/s, fraction of the measured HBM peak"""
        agg: dict = {}
        for name, recs in ops.KERNEL_TIMING.items():
            for tag, a, b in recs:
                agg.setdefault((name, tag), []).append(a.elapsed_time(b))
        ops.enable_kernel_timing([])
        rows = []
        for (name, tag), ts in agg.items():
            spec, mask = (tag, None) if name == "pml_tp_fwd" else (tag[0], tuple(bool(x) for x in tag[1:]))
            pl = next(p for p in plans._PLANS if getattr(p, "spec_id", None) == spec and hasattr(p, "NP") and hasattr(p, "DIN"))
            byts = edge_kernel_bytes(pl, n_atoms, n_edges, mask)
            avg = sum(ts) / len(ts)
            ach = byts / (avg * 1e-3) / 1e9
            rows.append({"kernel": "tp_fwd" if mask is None else "tp_bwd", "signature": pl.spec.key, "channels": pl.C,
                         "cotangents": None if mask is None else "".join(c for c, m in zip("hYR", mask) if m),
                         "launches_per_step": len(ts) / steps, "avg_launch_ms": avg, "ms_per_step": sum(ts) / steps,
                         "algorithmic_bytes_per_launch": byts, "achieved_gbs": ach, "frac": ach / peak})
        rows.sort(key=lambda r: -r["ms_per_step"])
        return rows

    def run_config(name: str, steps: int, warmup: int, headline: bool) -> dict:
        cfg = wl.CONFIGS[name]
        training = cfg["kind"] == "training"
        data, targets, _, _ = wl.make_batch(name, rank)
        model = wl.build_model(name).to(dev)
        if not training:
            model.eval()
        res: dict = {"kind": cfg["kind"], "family": cfg["family"]}
        d_dev = {k: (v.to(dev) if torch.is_tensor(v) and v.dim() > 0 else v) for k, v in data.items()}
        nl_ms = None
        if "edge_index" not in d_dev:  # the water box: edges come from the GPU neighbour list (timed by itself)
            from physicsml_b200.neighbour_list import radius_graph

            rc = cfg["model"]["cut_off"]
            for _ in range(2):
                ei, sh, plan = radius_graph(d_dev["coordinates"], rc, pbc=(True, True, True), cell=d_dev["cell"])
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                ei, sh, plan = radius_graph(d_dev["coordinates"], rc, pbc=(True, True, True), cell=d_dev["cell"])
            torch.cuda.synchronize()
            nl_ms = (time.perf_counter() - t0) / 5 * 1e3
            d_dev.update(edge_index=ei, cell_shift_vector=sh)
            data = dict(data, edge_index=ei.cpu(), cell_shift_vector=sh.cpu())
        t_dev = {k: v.to(dev) for k, v in targets.items()}
        n_atoms, n_edges = int(d_dev["coordinates"].shape[0]), int(d_dev["edge_index"].shape[1])
        host = {k: (v.pin_memory() if torch.is_tensor(v) and v.dim() > 0 else v) for k, v in data.items()}
        host_t = {k: v.pin_memory() for k, v in targets.items()}
        h2d = sum(v.numel() * v.element_size() for v in host.values() if torch.is_tensor(v) and v.dim() > 0)
        use_graph = (not args.eager) and n_atoms <= 6000  # large single structures are not launch-bound

        if training:
            h2d += sum(v.numel() * v.element_size() for v in host_t.values())
            reducer = parallel.GradientReducer([p for p in model.parameters() if p.requires_grad and p.numel() > 0]) if world > 1 else None
            opt = wl.make_optimizer(model, grad_scale=1.0 / world)
            if reducer is not None:
                reducer.attach(opt)

            def step(d, t):
                opt.zero_grad(set_to_none=True)
                loss = wl.training_loss(model, d, t)
                loss.backward()
                if reducer is not None:
                    reducer.finish()
                opt.step()
                return loss.detach()

            graphed = None
            if use_graph:
                try:
                    graphed = GraphedStep(model, opt, d_dev, t_dev, wl.training_loss,
                                          after_backward=(reducer.finish if reducer is not None else None), warmup=3)
                    keep_alive.append(graphed)
                except Exception as exc:  # capture is an optimisation: report and continue eagerly
                    print(f"[bench] {name}: CUDA-graph capture failed ({type(exc).__name__}: {exc}); running eagerly", file=sys.stderr)
                    graphed = None

            def resident():
                return graphed() if graphed is not None else step(d_dev, t_dev)

            out_host = torch.zeros(1).pin_memory()

            def e2e():
                if graphed is not None:
                    graphed.load(host, host_t)  # pinned host -> static device buffers
                    loss = graphed()
                else:
                    d = {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) and v.dim() > 0 else v) for k, v in host.items()}
                    t = {k: v.to(dev, non_blocking=True) for k, v in host_t.items()}
                    loss = step(d, t)
                out_host.copy_(loss.reshape(1), non_blocking=True)
                torch.cuda.current_stream().synchronize()  # the caller reads the loss every step
                return float(out_host[0])

            d2h = 4

            def eager():
                return step(d_dev, t_dev)
        else:
            graphed = None
            if use_graph:
                try:
                    graphed = GraphedEval(model, d_dev, wl.energy_forces, warmup=3)
                    keep_alive.append(graphed)
                except Exception as exc:
                    print(f"[bench] {name}: CUDA-graph capture failed ({type(exc).__name__}: {exc}); running eagerly", file=sys.stderr)
                    graphed = None
            d_plain = dict(d_dev)
            if nl_ms is not None:
                d_plain["_edge_plan"] = plan

            def resident():
                return graphed() if graphed is not None else wl.energy_forces(model, d_plain)

            G = int(data["num_graphs"])
            md_box: list = []
            e_host = torch.zeros(G, 1).pin_memory()
            f_host = torch.zeros(n_atoms, 3).pin_memory()

            def e2e():
                if graphed is not None:
                    graphed.load(host)
                    e, f = graphed()
                elif nl_ms is not None:
                    # single structure from host coordinates through the public per-step API (md.MDCalculator: copy the
                    # coordinates in, rebuild the neighbour list on the device, evaluate energy and forces)
                    if not md_box:
                        from physicsml_b200.md import MDCalculator

                        md_box.append(MDCalculator(model, cfg["model"]["cut_off"], data["node_attrs"], cell=data["cell"]))
                    e, f = md_box[0].energy_forces(host["coordinates"])
                else:
                    d = {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) and v.dim() > 0 else v) for k, v in host.items()}
                    e, f = wl.energy_forces(model, d)
                e_host.copy_(e, non_blocking=True)
                f_host.copy_(f, non_blocking=True)
                torch.cuda.current_stream().synchronize()
                return e_host

            d2h = 4 * G + 12 * n_atoms
            if nl_ms is not None:  # the e2e leg of the box ships the coordinates; species and cell live in the calculator, edges are built on the device
                h2d = host["coordinates"].numel() * host["coordinates"].element_size()

            def eager():
                return wl.energy_forces(model, d_plain)

        sampler = ClockSampler(local) if (rank == 0 and headline) else None
        ms_dev, launches = timed(resident, steps, warmup)
        if headline:
            # nvidia-smi delivers its first line ~100 ms after it starts and the timed region is ~100 ms long: keep the
            # SAME step running, untimed, for another 0.6 s so that the clock / throttle samples are taken under its load
            # (every rank runs the same number of extra steps: the step holds a collective)
            for _ in range(int(600.0 / max(ms_dev, 0.05)) + 1):
                resident()
            torch.cuda.synchronize()
        if sampler is not None:
            res["clocks"] = sampler.stop()
            res["clocks"]["window"] = "timed steps + 0.6 s of the same step, untimed"
        # a replayed graph cannot carry per-kernel events: time the fused edge kernels in an eager pass of the SAME step
        # (same inputs, same kernels), and count the kernels of one step the same way
        # ... with the side stream off, so that every launch runs alone between its two events (in the timed step the radial
        # MLPs run concurrently on a second stream; a launch bracketed there would be charged for its neighbours' SM time)
        from physicsml_b200 import mace as _mace

        side_was, _mace.SIDE_STREAM[0] = _mace.SIDE_STREAM[0], False
        try:
            _, launches = timed(eager, min(steps, 10), 2, kernel_timing=True)
        finally:
            _mace.SIDE_STREAM[0] = side_was
        table = edge_kernel_table(n_atoms, n_edges, min(steps, 10))
        ms_e2e = None
        if not args.no_e2e:
            ms_e2e, _ = timed(e2e, steps, warmup)
        total_atoms = float(n_atoms)
        if world > 1:
            t = torch.tensor([float(n_atoms), float(n_edges)], device=dev)
            per_rank = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(per_rank, t)
            total_atoms = float(sum(x[0] for x in per_rank))
            res["per_rank"] = [{"atoms": int(x[0]), "edges": int(x[1])} for x in per_rank]
            if training and reducer is not None:
                res["allreduce"] = reducer.describe()
        res.update({"ms_per_step": ms_dev, "atoms_per_s": total_atoms / (ms_dev * 1e-3), "atoms": n_atoms, "edges": n_edges,
                    "total_atoms": total_atoms, "cuda_graph": graphed is not None, "gpu_launches": launches,
                    "e2e_ms_per_step": ms_e2e, "e2e_atoms_per_s": (total_atoms / (ms_e2e * 1e-3)) if ms_e2e else None,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "edge_kernels": table})
        if nl_ms is not None:
            res["neighbour_list_ms"] = nl_ms
            res["atoms_per_s_with_neighbour_list"] = total_atoms / ((ms_dev + nl_ms) * 1e-3)
            res["md_loop"] = md_loop_rate(model, cfg, data, dev, steps)
        if headline and rank == 0 and not args.no_cpu_baseline and world == 1:
            res["parity"] = parity_vs_oracle(name, model, data, targets, dev)
        return res

    def md_loop_rate(model, cfg, data, dev, steps):
        """the consumer of this config: an MD-style loop through md.MDCalculator — host coordinates in, energy
        and forces back on the host EVERY step, atoms moving ~0.03 A per step, the neighbour list rebuilt on the GPU as
        MDCalculator decides (every step at this size; the reference also rebuilds it every step,
        plugins/openmm/openmm_graph.py:148-156, with its 1.37 s linked cell)"""
        from physicsml_b200.md import MDCalculator

        md = MDCalculator(model, cfg["model"]["cut_off"], data["node_attrs"], cell=data["cell"])  # 24 k atoms: eager, exact list every step
        n = int(data["coordinates"].shape[0])
        gen = torch.Generator().manual_seed(5)
        pos = data["coordinates"].clone()
        host_pos = torch.empty(n, 3).pin_memory()
        f_host = torch.empty(n, 3).pin_memory()
        e_host = torch.empty(1).pin_memory()
        times = []
        mallocs = 0
        try:
            for i in range(steps + 3):
                pos = pos + 0.017 * torch.randn(n, 3, generator=gen)  # |dr| ~ 0.03 A per step
                host_pos.copy_(pos)
                torch.cuda.synchronize()
                if i == 3:
                    mallocs = torch.cuda.memory_stats()["num_device_alloc"]
                t0 = time.perf_counter()
                e, f = md.energy_forces(host_pos)
                e_host.copy_(e, non_blocking=True)
                f_host.copy_(f, non_blocking=True)
                torch.cuda.current_stream().synchronize()
                if i >= 3:
                    times.append(time.perf_counter() - t0)
            mallocs = torch.cuda.memory_stats()["num_device_alloc"] - mallocs
        finally:
            # MDCalculator switched the caching allocator to expandable segments (a list of a different length every
            # step; see md.py): back to the default for the configs that follow
            from physicsml_b200.md import set_expandable_segments

            set_expandable_segments(False)
        ms = 1e3 * sum(times) / len(times)
        return {"ms_per_step": ms, "ms_per_step_max": 1e3 * max(times), "atoms_per_s": n / (ms * 1e-3), "steps": len(times),
                "neighbour_list_builds": md.n_builds, "skin_A": md.skin, "edges_in_list": md.n_edges, "cuda_mallocs_in_timed_steps": mallocs,
                "timing": "host wall clock per step, coordinates from / results to pinned host memory"}

    def parity_vs_oracle(name, model, data, targets, dev):
        """step-0 check inside the benchmark run: the GPU model vs the CPU oracle (same weights) on the first
        PARITY_MOLECULES molecules of the timed batch: energies, forces and the training loss"""
        ref, kind = _cpu_model(name)
        sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        ref.load_state_dict(sd, strict=False)
        n_mol = PARITY_MOLECULES
        keep = data["batch"] < n_mol
        n = int(keep.sum())
        em = data["edge_index"][0] < n
        sub = {"coordinates": data["coordinates"][:n], "node_attrs": data["node_attrs"][:n], "batch": data["batch"][:n],
               "num_graphs": torch.tensor(n_mol), "edge_index": data["edge_index"][:, em]}
        sub_t = {"energy": targets["energy"][:n_mol], "forces": targets["forces"][:n]}
        e_ref, f_ref = wl.energy_forces(ref, sub)
        l_ref = wl.training_loss(ref, sub, sub_t)
        sub_d = {k: (v.to(dev) if torch.is_tensor(v) and v.dim() > 0 else v) for k, v in sub.items()}
        e, f = wl.energy_forces(model, sub_d)
        l = wl.training_loss(model, sub_d, {k: v.to(dev) for k, v in sub_t.items()})
        scale = e_ref.abs().max().item()
        return {"against": kind, "molecules": n_mol, "atoms": n, "energy_rel": (e.cpu() - e_ref).abs().max().item() / scale,
                "force_abs": (f.cpu() - f_ref).abs().max().item(), "loss_rel": abs(l.item() - l_ref.item()) / abs(l_ref.item()),
                "tolerance": {"energy_rel": 1e-5, "force_abs": 1e-4}}

    def shutdown():
        """Leave the process group without hanging: the captured steps hold NCCL work, so drop the graphs first; a
        watchdog forces the exit if the communicator teardown still blocks (the JSON line is already flushed)."""
        if world == 1:
            return
        import gc
        import threading

        sys.stdout.flush()
        threading.Timer(30.0, lambda: os._exit(0)).start()
        keep_alive.clear()
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
        os._exit(0)

    head = run_config(WORKLOAD, args.steps, args.warmup, headline=True)
    others: dict = {}
    wanted = OTHERS if args.configs == "all" else ([] if args.configs == "headline" else args.configs.split(","))
    for name in wanted:
        if world > 1 and wl.CONFIGS[name]["kind"] != "training":
            continue  # single-structure / batched inference is reported at one GPU (replicas only)
        try:
            r = run_config(name, min(args.steps, 10), 3, headline=False)
        except Exception as exc:  # never lose the headline line to a secondary workload
            r = {"error": f"{type(exc).__name__}: {exc}"}
            if world > 1:
                raise
        others[name] = r
        torch.cuda.empty_cache()

    if rank != 0:
        shutdown()
        return

    # headline roofline: the fused edge kernel that takes the most time of the step (forward or cotangent)
    roof = None
    if head["edge_kernels"]:
        top = head["edge_kernels"][0]
        traffic, tsrc = None, None
        tpath = os.path.join(ROOT, "profiles", "k4_dram_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                rec = json.load(f).get(f"{WORKLOAD}:{top['kernel']}:{top['signature']}" + (f":{top['cotangents']}" if top["cotangents"] else ""))
            if rec:
                traffic, tsrc = rec["dram_bytes_per_launch"], rec["source"]
        roof = {"bound": "hbm", "achieved": top["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": top["frac"],
                "traffic": traffic, "traffic_source": tsrc,
                "kernel": f"{top['kernel']}<{top['signature']}>" + (f" cotangents {top['cotangents']}" if top["cotangents"] else ""),
                "algorithmic_bytes_per_launch": top["algorithmic_bytes_per_launch"], "avg_launch_ms": top["avg_launch_ms"],
                "launches_per_step": top["launches_per_step"], "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({which})",
                "all_edge_kernels": [{k: r[k] for k in ("kernel", "signature", "cotangents", "launches_per_step", "avg_launch_ms", "frac")}
                                     for r in head["edge_kernels"]]}

    cpu = None
    if not args.no_cpu_baseline and world == 1:  # reported on rank 0 at N = 1 only, on a bounded sample
        rate, cores, sample, _, kind = cpu_reference_step_rate(args.cpu_baseline_molecules, 2, 1)
        cpu = {"value": rate, "unit": "atoms/s", "cores": cores, "kind": kind, "sample": sample}

    # informational: the same training step in plain eager PyTorch ON THE SAME GPU (the reference's own modules over the
    # e3nn shims, stock autograd, torch.optim.Adam) — what a user of the reference gets by calling .cuda() today
    gpu_eager = None
    if not args.no_cpu_baseline and world == 1:
        try:
            gpu_eager = gpu_eager_step_rate(dev)
        except Exception as exc:
            gpu_eager = {"error": f"{type(exc).__name__}: {exc}"}

    config = shared_config(world)
    line = {
        "metric": METRIC, "value": head["atoms_per_s"], "unit": "atoms/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
        "parallelism": f"dp{world}",
        "l2": "flushed between steps (256 MiB memset outside the per-step event pair)",
        "e2e": {"value": head["e2e_atoms_per_s"], "unit": "atoms/s", "h2d_bytes_per_step": head["h2d_bytes_per_step"],
                "d2h_bytes_per_step": head["d2h_bytes_per_step"], "ms_per_step": head["e2e_ms_per_step"]},
        "gpu_launches": head["gpu_launches"] * args.steps, "gpu_launches_per_step": head["gpu_launches"],
        "cuda_graph": head["cuda_graph"], "clocks": head.get("clocks"), "roofline": roof, "cpu_baseline": cpu,
        "parity": head.get("parity"), "per_rank": head.get("per_rank"), "allreduce": head.get("allreduce"),
        "gpu_eager_baseline": gpu_eager,
        "other_configs": others,
    }
    print(json.dumps(line), flush=True)
    shutdown()


if __name__ == "__main__":
    main()
