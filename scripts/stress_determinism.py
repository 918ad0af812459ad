"""Repeat the forward ops of a small model on fixed inputs and report any run-to-run difference (all forward kernels are
meant to be deterministic: no atomics on the forward path)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import helpers  # noqa: E402
from physicsml_b200 import _lib, ops  # noqa: E402

POISON = os.environ.get("PML_POISON_SMEM", "1") == "1"  # NaN the shared memory of every SM before each kernel
from physicsml_b200.mace import PooledMACEModule  # noqa: E402

dev = torch.device("cuda:0")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
kw = dict(cut_off=5.0, num_bessel=8, num_polynomial_cutoff=6, max_ell=2, num_interactions=2, num_node_feats=4, num_edge_feats=0,
          hidden_irreps="4x0e + 4x1o", avg_num_neighbours=10.0, correlation=2)
torch.manual_seed(0)
model = PooledMACEModule(mlp_irreps="5x0e", **kw).to(dev)
pos, z, batch = helpers.random_molecules(10, 10, 22, 4, seed=101)
data = helpers.to_device(helpers.make_batch(pos, z, batch, 4, 5.0), dev)

# record every custom-op call of one forward, then replay each op many times on its recorded inputs
calls = []
orig = ops._call


def rec(name, *args, **kw_):
    calls.append(name)
    return orig(name, *args, **kw_)


outs = {}
with torch.no_grad():
    e0 = model(dict(data))["y_graph_scalars"].clone()
bad = 0
for i in range(reps):
    # poison the caching allocator's free blocks: a kernel that leaves part of an output unwritten shows up
    junk = [torch.full((n,), float("nan"), device=dev) for n in (1 << 20, 1 << 18, 1 << 16, 4096, 777)]
    del junk
    with torch.no_grad():
        if POISON:
            _orig_call = ops._call

            def poisoned(name, *a, **k):
                _lib.lib().pml_debug_poison_smem(torch.cuda.current_stream().cuda_stream)
                return _orig_call(name, *a, **k)

            ops._call = poisoned
        try:
            e = model(dict(data))["y_graph_scalars"]
        finally:
            if POISON:
                ops._call = _orig_call
    d = (e - e0).abs().max().item()
    if not (d <= 1e-7):
        bad += 1
        print("run", i, "energy differs: max abs", d, flush=True)
print("whole forward (poisoned allocator, tolerance 1e-7):", bad, "of", reps, "runs differ")

# per-op bisection with hooks on the torch.library ops
import physicsml_b200.mace as M  # noqa: E402

inter = model.mace.interactions[0]
plan = ops.EdgePlan(data["edge_index"], pos.shape[0])
re_ = model.mace.radial_embedding
with torch.no_grad():
    Y, B = ops.edge_featurise(data["coordinates"], plan.snd, plan.rcv, None, None, re_.bessel_fn.bessel_weights, re_.r_max, re_.prefactor, 1.0, re_.p, 2)
    h = model.mace.node_embedding(data["node_attrs"])
    wh = inter.linear_node_feats(h)
    R = inter.net_R_channels(B)
    A = ops.tp_scatter(wh, Y, R, plan.gather, plan.rowptr, plan.perm, inter.tp.handle)

    def check(name, fn, ref):
        n = 0
        for _ in range(reps):
            if not torch.equal(fn(), ref):
                n += 1
        print(f"{name}: {n} of {reps} differ", flush=True)

    check("edge_featurise", lambda: ops.edge_featurise(data["coordinates"], plan.snd, plan.rcv, None, None, re_.bessel_fn.bessel_weights, re_.r_max, re_.prefactor, 1.0, re_.p, 2)[0], Y)
    check("radial_mlp (tc gemm)", lambda: inter.net_R_channels(B), R)
    check("tp_scatter", lambda: ops.tp_scatter(wh, Y, R, plan.gather, plan.rowptr, plan.perm, inter.tp.handle), A)
    A2 = inter.linear(A)
    check("linear", lambda: inter.linear(A), A2)
    A3 = inter.mix_layer(A2, data["node_attrs"])
    check("mix_layer", lambda: inter.mix_layer(A2, data["node_attrs"]), A3)
    m = model.mace.messages[0](A3, data["node_attrs"])
    check("sym_contract", lambda: model.mace.messages[0](A3, data["node_attrs"]), m)
