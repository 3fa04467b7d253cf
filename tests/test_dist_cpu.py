"""World-size-2 tests of the data-parallel host logic on CPU (gloo): the flat-gradient bucket exchange that
Trainer captures into its CUDA graph, and the batch sharding of generation.  (The kernels themselves are
rank-local: the path has exactly one collective, the gradient all-reduce.)"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mamba_b200 import train
        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Tanh(), torch.nn.Linear(16, 4))
        fg = train.FlatGrads(model.parameters(), bucket_mb=1e-4)   # tiny buckets: several all-reduces
        assert len(fg.buckets) >= 3
        assert all(p.grad.data_ptr() >= fg.flat.data_ptr() for p in model.parameters())
        g = torch.Generator().manual_seed(100)
        x = torch.randn(6, 8, generator=g)
        y = torch.randn(6, 4, generator=g)
        lo, hi = train.shard_rows(6, rank, world)
        fg.zero()
        torch.nn.functional.mse_loss(model(x[lo:hi]), y[lo:hi]).backward()
        fg.allreduce_mean(world)
        eager = fg.flat.clone()
        # same exchange launched bucket by bucket from the backward hooks (the overlapped form Trainer uses)
        fg.overlap_with_backward(world)
        fg.zero()
        torch.nn.functional.mse_loss(model(x[lo:hi]), y[lo:hi]).backward()
        fg.finish()
        assert torch.equal(fg.flat, eager)
        assert all(n == 0 for n in fg._pending)
        q.put((rank, eager, (lo, hi)))
    finally:
        dist.destroy_process_group()


def test_flat_grad_allreduce_equals_single_process_mean():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(world)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][2] == (0, 3) and res[1][2] == (3, 6)
    assert torch.equal(res[0][1], res[1][1])      # every rank holds the same reduced gradient
    # reference: one process, both shards, mean of the per-shard gradients (== DDP semantics)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Tanh(), torch.nn.Linear(16, 4))
    g = torch.Generator().manual_seed(100)
    x = torch.randn(6, 8, generator=g)
    y = torch.randn(6, 4, generator=g)
    grads = []
    for lo, hi in ((0, 3), (3, 6)):
        model.zero_grad()
        torch.nn.functional.mse_loss(model(x[lo:hi]), y[lo:hi]).backward()
        grads.append(torch.cat([p.grad.flatten() for p in model.parameters()]))
    want = (grads[0] + grads[1]) / 2
    assert torch.allclose(res[0][1], want, rtol=1e-6, atol=1e-7)


def test_shard_rows_partitions_exactly():
    from mamba_b200 import train
    for n in (1, 5, 10, 17):
        for w in (1, 2, 4, 8):
            parts = [train.shard_rows(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1
