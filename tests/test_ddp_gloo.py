"""world_size-2 check of the gradient all-reduce plumbing on the gloo backend (CPU)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers  # noqa: F401


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from physicsml_b200.parallel import allreduce_gradients, shard_indices

    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(2, 3)), torch.nn.Parameter(torch.zeros(0))]
    for i, p in enumerate(params):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    allreduce_gradients(params)
    ok = all(torch.allclose(p.grad, torch.full_like(p, 1.5 * (i + 1))) for i, p in enumerate(params))
    # a parameter whose gradient is missing on ONE rank (an element absent from that rank's batch): identical bucket layout
    # on both ranks, zeros substituted, every rank ends with the average
    for i, p in enumerate(params):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    if rank == 1:
        params[0].grad = None
    allreduce_gradients(params)
    ok = ok and torch.allclose(params[0].grad, torch.full_like(params[0], 0.5)) and torch.allclose(
        params[1].grad, torch.full_like(params[1], 3.0))
    # the bucketed reducer driven by autograd hooks: two buckets, reduction started from the hooks from the second step on
    from physicsml_b200.parallel import GradientReducer

    torch.manual_seed(1)
    ws = [torch.nn.Parameter(torch.randn(6, 6)) for _ in range(4)]
    red = GradientReducer(ws, n_buckets=2)
    for step in range(3):
        for w in ws:
            w.grad = None
        x = torch.full((1, 6), float(rank + 1))
        for w in ws:
            x = x @ w
        x.sum().backward()
        local = [w.grad.clone() for w in ws]
        red.finish()
        red.write_back()
        gathered = [[torch.zeros_like(g) for _ in range(world)] for g in local]
        for g, lst in zip(local, gathered):
            dist.all_gather(lst, g)
        ok = ok and all(torch.allclose(w.grad, sum(lst) / world, rtol=1e-6, atol=1e-6) for w, lst in zip(ws, gathered))
    ok = ok and red.launched_early >= 2 and sorted(red.order) == [0, 1, 2, 3] and red.order[0] == 3
    shard = list(shard_indices(7, rank, world))
    out[rank] = (ok, shard)
    dist.destroy_process_group()


def test_flat_bucket_allreduce_two_ranks():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0][0] and out[1][0]
    assert sorted(out[0][1] + out[1][1]) == list(range(7))
