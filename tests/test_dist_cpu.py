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


class _FakeNorm(torch.nn.Module):
    """Stand-in for models.mamba.RMSNorm's fused form: forward(x, residual) -> (norm(x + residual), x + residual)."""

    def __init__(self, d):
        super().__init__()
        self.weight = torch.nn.Parameter(torch.ones(d))

    def forward(self, x, residual):
        r = residual if x is None else x + residual
        return r * torch.rsqrt(r.square().mean(-1, keepdim=True) + 1e-5) * self.weight, r


class _FakeLayer(torch.nn.Module):
    def __init__(self, d):
        super().__init__()
        self.norm = _FakeNorm(d)
        self.mixer = torch.nn.Sequential(torch.nn.Linear(d, d), torch.nn.Tanh())


class _FakeMamba(torch.nn.Module):
    """Layout-P skeleton (embedding tied to lm_head, pre-norm residual layers, norm_f) in plain torch: what Trainer's
    staged backward walks, without the CUDA-only mixer."""
    layout = "P"

    def __init__(self, d=8, n_layers=4):
        super().__init__()
        from mamba_b200.configs import common as cc
        self.embedding = torch.nn.Embedding(cc.vocab_size, d)
        self.metadata_embedding = torch.nn.Embedding(cc.metadata_vocab_size, d)
        self.layers = torch.nn.ModuleList(_FakeLayer(d) for _ in range(n_layers))
        self.norm_f = _FakeNorm(d)
        self.lm_head = torch.nn.Linear(d, cc.vocab_size, bias=False)
        self.lm_head.weight = self.embedding.weight

    def _embed(self, tokens, meta):
        return torch.cat((self.metadata_embedding(meta), self.embedding(tokens)), dim=-2)

    def _head(self, x):
        return self.lm_head(x)

    def forward(self, tokens, meta):
        resid, hidden = self._embed(tokens, meta), None
        for layer in self.layers:
            normed, resid = layer.norm(hidden, resid)
            hidden = layer.mixer(normed)
        normed, _ = self.norm_f(hidden, resid)
        return self._head(normed[:, meta.shape[-1]:])


def _staged_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mamba_b200 import synthetic, train
        train.loss_fn = train.loss_fn_torch          # the fused loss is CUDA-only; same function in torch ops
        out = []
        for stages in (1, 2, 4):
            # every rank builds DIFFERENT initial weights: Trainer must broadcast rank 0's (DDP semantics,
            # train_parallel.py:151), otherwise the ranks would not agree at the end
            torch.manual_seed(7 * rank)
            model = _FakeMamba()
            tr = train.Trainer(model, lr=1e-2, autocast_dtype=None, world_size=world, batch_size=2, block_len=12,
                               use_graph=False, stages=stages)
            assert (tr._stage_groups is None) == (stages == 1)
            if stages > 1:
                assert len(tr.grads.buckets) == stages + 1   # the layer groups + the embedding / tied-head stage
                assert all(p.grad is not None for p in model.parameters())
            losses = []
            for i in range(3):
                src, trg, meta = synthetic.batch(2, 12, seed=50 + 10 * i + rank)
                losses.append(float(tr.step(src, trg, meta)))
            out.append((losses, torch.cat([p.detach().flatten() for p in model.parameters()]).numpy()))  # by value
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_staged_backward_with_per_stage_allreduce_equals_single_exchange():
    """Trainer's staged backward (graph cut into layer groups, bucket i all-reduced as soon as group i's gradients
    exist) trains exactly like one exchange after a plain backward, and every rank ends with the same parameters."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_staged_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=300) for _ in range(world)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank_out in res:
        (l1, p1), (l2, p2), (l4, p4) = rank_out[1]
        assert l1 == l2 == l4, (l1, l2, l4)
        assert (p1 == p2).all() and (p1 == p4).all()
    assert (res[0][1][0][1] == res[1][1][0][1]).all()   # ranks agree
