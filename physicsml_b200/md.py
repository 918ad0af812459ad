"""Single-structure energy + forces loop for molecular dynamics — the real consumer of BASELINE
configs[3].  Mirrors what the reference's OpenMM graph module does per MD step
(``plugins/openmm/openmm_graph.py:119-176``: scale positions, override the cell from the box vectors, build the
neighbour list, run the module, scale the output) and what its ASE calculator returns (``plugins/ase/ase_graph.py:17-64``:
energy and forces), with three differences that matter at 24 000 atoms:

* **Verlet skin** (default on when the evaluation is graph-captured, off for large eager systems, where the 0.7 ms GPU
  list is simply rebuilt every step): the neighbour list is built with ``cut_off + skin`` and reused until some atom has
  moved more than ``skin / 2`` since the build (one tiny kernel + one 4-byte read-back per step).  Results are EXACT, not approximate: an
  edge longer than the model's cut-off contributes exactly zero to every layer — the polynomial cut-off makes its Bessel
  features 0, the bias-free radial MLP maps 0 to 0 (SiLU(0) = 0), so its tensor-product weights vanish.
* **No per-step host work** between rebuilds: the evaluation is captured in a CUDA graph per neighbour list (optional; on
  by default below ``graph_max_atoms``) and replayed; positions are copied into the static buffer.
* **Allocator**: an eagerly evaluated list has a slightly different length every step, so every edge-sized tensor (up
  to 10 GB at 2 M edges) changes size every step.  PyTorch's default caching allocator cannot reuse a cached block for
  a larger request: measured on the 24 k-atom box, two fresh ~10 GB cudaMallocs per step (+19 GiB reserved per step)
  until the device is full, then an 870 ms stall while everything is freed (scripts/md_probe.py).  The calculator
  therefore switches the allocator to expandable segments (``expandable_segments=True``; a process-wide PyTorch
  allocator setting) and, after the first step, maps a 10 % margin once (one block larger than everything reserved so far,
  freed at once): the same loop then runs without a single cudaMalloc, 32.4-34 ms per step as the list grows.
* **Analytic forces**: -dE/dx comes from the fused cotangent kernels in the same call (the OpenMM route differentiates
  the TorchScript module from outside).

The reference rebuilds the list on every step on the device the module lives on; its vendored linked cell takes 1.37 s
for the 24 k-atom box on 8 CPU threads — the number the ``neighbour_list_ms`` of this loop compares with.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from .graphs import GraphedEval
from .neighbour_list import radius_graph
from .ops import _call, _p, _stream


def set_expandable_segments(on: bool) -> None:
    """process-wide CUDA caching-allocator setting (the runtime form of PYTORCH_CUDA_ALLOC_CONF=expandable_segments:...)"""
    setting = f"expandable_segments:{'True' if on else 'False'}"
    fn = getattr(torch._C, "_accelerator_setAllocatorSettings", None)
    if fn is not None:
        fn(setting)
    else:
        torch.cuda.memory._set_allocator_settings(setting)


class MDCalculator:
    def __init__(self, module: torch.nn.Module, cut_off: float, node_attrs: torch.Tensor, cell: Optional[torch.Tensor] = None,
                 skin: Optional[float] = None, position_scaling: float = 1.0, output_scaling: float = 1.0, self_interaction: bool = False,
                 use_graph: Optional[bool] = None, graph_max_atoms: int = 6000, y_output: str = "y_graph_scalars",
                 total_atomic_energy: Optional[float] = None, expandable_segments: Optional[bool] = None) -> None:
        """module: a ``forward(data) -> {"y_graph_scalars": ...}`` module on a CUDA device (``PooledMACEModule`` /
        ``PooledNequipModule`` or a patched reference module); node_attrs [N, Z] one-hot species (+ features);
        cell [3,3] or None (open boundaries); position_scaling / output_scaling as in the reference's OpenMM module
        (nm -> Angstrom, eV -> kJ/mol).  expandable_segments: let the CUDA caching allocator grow its segments in place
        (default: on for eagerly evaluated systems, see the module docstring; False leaves the allocator alone)."""
        self.module = module.eval()
        self.dev = next(module.parameters()).device
        if self.dev.type != "cuda":
            raise RuntimeError("MDCalculator needs the module on a CUDA device (no CPU fallback)")
        self.cut_off = float(cut_off)
        self.position_scaling, self.output_scaling = float(position_scaling), float(output_scaling)
        self.self_interaction, self.y_output = self_interaction, y_output
        self.node_attrs = node_attrs.to(self.dev, torch.float32).contiguous()
        self.N = self.node_attrs.shape[0]
        self.cell = None if cell is None else cell.to(self.dev, torch.float32).reshape(3, 3).contiguous()
        self._cell_host = None if cell is None else cell.detach().cpu().to(torch.float32).reshape(3, 3).clone()
        self.use_graph = (self.N <= graph_max_atoms) if use_graph is None else bool(use_graph)
        # Skin default.  A list built with cut_off + skin has (1 + skin / cut_off)^3 times the edges (+27 % at 6 + 0.5 A), all of
        # which the edge kernels still process; the GPU neighbour list itself costs 0.7 ms at 24 000 atoms.  So the skin pays
        # where a rebuild is expensive — a captured graph has to be re-captured (small systems) — and NOT on large eager
        # systems, which simply rebuild the exact list every step.
        self.skin = float(skin) if skin is not None else (0.5 if self.use_graph else 0.0)
        self.total_atomic_energy = total_atomic_energy
        if expandable_segments is None:
            expandable_segments = not self.use_graph
        if expandable_segments:
            set_expandable_segments(True)
        self._headroom = 0.1 if expandable_segments else 0.0  # fraction of the reserved memory mapped as a margin after step 1
        self.pos = torch.zeros(self.N, 3, device=self.dev)       # static input buffer
        self.ref_pos = torch.zeros(self.N, 3, device=self.dev)   # positions the current list was built from
        self._disp = torch.zeros(1, device=self.dev)
        self._disp_host = torch.zeros(1).pin_memory()
        self._data: Optional[dict] = None
        self._graph: Optional[GraphedEval] = None
        self._arena = ops.ZeroArena()  # shared by the successive captures (one per neighbour list)
        self.n_builds = 0
        self.n_steps = 0
        self.n_edges = 0

    # ------------------------------------------------------------------------------------------------------------------
    def _needs_rebuild(self) -> bool:
        if self._data is None or self.skin <= 0.0:
            return True
        self._disp.zero_()
        _call("pml_max_disp2", _p(self.pos), _p(self.ref_pos), self.N, _p(self._disp), _stream())
        self._disp_host.copy_(self._disp, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # 4 bytes; the previous step's forces were already read by the caller
        return float(self._disp_host[0]) > (0.5 * self.skin) ** 2

    def _rebuild(self) -> None:
        if self._graph is not None:
            self._graph.release()  # frees the static buffers of the previous list before the new ones are allocated
        self._graph = None
        self._data = None
        pbc = (True, True, True) if self.cell is not None else None
        ei, sh, plan = radius_graph(self.pos, self.cut_off + self.skin, pbc=pbc, cell=self.cell, self_interaction=self.self_interaction)
        self.ref_pos.copy_(self.pos)
        data = {"coordinates": self.pos, "edge_index": ei, "node_attrs": self.node_attrs,
                "batch": torch.zeros(self.N, dtype=torch.long, device=self.dev), "num_graphs": torch.tensor(1)}
        if self.cell is not None:
            data["cell"], data["cell_shift_vector"] = self.cell, sh
        if self.total_atomic_energy is not None:
            data["total_atomic_energy"] = torch.tensor([float(self.total_atomic_energy)], device=self.dev)
        plan._src = None  # this loop manages staleness itself: the plan is replaced together with the edges
        self._data, self._plan = data, plan
        self.n_builds += 1
        self.n_edges = int(ei.shape[1])
        if self.use_graph:
            self._graph = GraphedEval(self.module, data, self._evaluate, warmup=2, arena=self._arena)

    def _evaluate(self, module, data):
        from .mace import compute_forces_by_gradient

        d = dict(data)
        pos = d["coordinates"].detach().requires_grad_(True)
        d["coordinates"] = pos
        d["_edge_plan"] = self._plan  # int32 edge copies / CSR of the current list (persistent buffers: graph-safe)
        e = module(d)[self.y_output]
        f = compute_forces_by_gradient(e, pos, training=False)
        return e.reshape(-1)[:1].clone(), f  # clone: the pooled energy may live in the zero arena, which the next capture reuses

    # ------------------------------------------------------------------------------------------------------------------
    def energy_forces(self, positions: torch.Tensor, boxvectors: Optional[torch.Tensor] = None):
        """one MD step: -> (energy [1] , forces [N, 3]) on the device, in the caller's units (output_scaling applied;
        forces additionally multiplied by position_scaling: d/d(caller position) = scaling * d/d(model position))"""
        p = positions.to(self.dev, torch.float32, non_blocking=True)
        if self.position_scaling != 1.0:
            p = p * self.position_scaling
        force_rebuild = False
        if boxvectors is not None:
            b = boxvectors.detach().to(torch.float32).reshape(3, 3)
            bh = (b.cpu() if b.is_cuda else b) * self.position_scaling
            if self._cell_host is None or not torch.equal(bh, self._cell_host):
                self._cell_host = bh.clone()
                self.cell = bh.to(self.dev).contiguous()
                force_rebuild = True
        self.pos.copy_(p, non_blocking=True)
        if force_rebuild or self._needs_rebuild():
            self._rebuild()
        if self._graph is not None:
            self._graph.data["coordinates"].copy_(self.pos, non_blocking=True)
            e, f = self._graph()
        else:
            e, f = self._evaluate(self.module, self._data)
        if self.n_steps == 0 and self._graph is None and self._headroom > 0.0:
            # Eager evaluation: the first step has just mapped the segment it needs.  Atoms move and the list grows by a
            # fraction of a percent per rebuild; map a margin once (and hand it straight back to the cache) so that later
            # steps never stall in a multi-GB cuMemMap (measured: one 50-80 ms step in ten on the 24 k-atom box).
            # (ONE block larger than everything reserved so far: the freed blocks of the step coalesce into one range of
            # the expandable segment, so the request extends the mapping instead of being served from the cache)
            try:
                margin = torch.empty(int((1.0 + self._headroom) * torch.cuda.memory_reserved(self.dev)), dtype=torch.uint8,
                                     device=self.dev)
                del margin
            except torch.OutOfMemoryError:
                pass  # a device that full has no room for a margin; the loop still works, with the occasional stall
        self.n_steps += 1
        s = self.output_scaling
        return (e * s if s != 1.0 else e), (f * (s * self.position_scaling) if s * self.position_scaling != 1.0 else f)

    def ase_results(self, positions) -> dict:
        """the two entries the reference's ASE calculator fills (ase_graph.py:58-61): python float energy, numpy forces"""
        e, f = self.energy_forces(torch.as_tensor(positions, dtype=torch.float32))
        return {"energy": float(e.cpu()[0]), "forces": f.detach().cpu().numpy()}
