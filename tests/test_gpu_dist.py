"""Two-GPU correctness of the data-parallel step over NCCL (skipped on a one-GPU box): what SCALE_rNN.json cannot
show — that N ranks compute the SAME update as one rank on the concatenated batch, that every
rank ends with identical parameters, and that the model also trains under torch's own DistributedDataParallel, the
wrapper the reference uses (train_parallel.py:143-185).

Tolerance (stated): fp32 end to end (autocast off); after three Adam steps the parameters of the 2-rank run and of the
1-rank run on the concatenated batch agree within 1e-3 of the largest parameter update on 99.9 % of all elements.
(Adam's first steps move a parameter by lr * sign(gradient): an element whose gradient is pure cancellation noise,
+1e-12 on one summation order and -1e-12 on the other, legitimately ends a full step apart; those few elements are
bounded by the largest possible difference, 2 * lr * steps.)"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _small_model():
    from mamba_b200.configs import common as cc
    from mamba_b200.models.mamba import Mamba, ModelArgs
    torch.manual_seed(0)
    return Mamba(ModelArgs(d_model=128, n_layer=5, vocab_size=cc.vocab_size, d_state=16, expand=2, d_conv=4,
                           pad_vocab_size_multiple=1, metadata_vocab_size=cc.metadata_vocab_size))


def _worker(rank, world, port, q, mode):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from mamba_b200 import synthetic, train
        model = _small_model().cuda()
        if rank == 1 and mode == "trainer":   # a rank that seeded differently must be overwritten by rank 0's weights
            with torch.no_grad():
                for p in model.parameters():
                    p.add_(0.01)
        src, trg, meta = synthetic.batch(4, 96, seed=5)
        lo, hi = train.shard_rows(4, rank, world)
        s, t, m = src[lo:hi].cuda(), trg[lo:hi].cuda(), meta[lo:hi].cuda()
        losses = []
        if mode == "trainer":
            tr = train.Trainer(model, lr=1e-3, autocast_dtype=None, world_size=world, batch_size=hi - lo, block_len=96)
            for _ in range(3):
                losses.append(float(tr.step(s, t, m)))
        else:   # the reference's own wrapper: DDP(model) + the python-launched step (train_parallel.py:151,173-185)
            ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[rank])
            opt = torch.optim.Adam(ddp.parameters(), lr=1e-3)
            for _ in range(3):
                losses.append(float(train.train_step(ddp, opt, s, t, m).detach()))
        torch.cuda.synchronize()
        flat = torch.cat([p.detach().flatten() for p in model.parameters()]).cpu().numpy()   # by value, not by fd
        q.put((rank, flat, losses))
        q.close()
        q.join_thread()
    finally:
        # no destroy_process_group: tearing NCCL down under a live CUDA graph that holds captured all-reduces can
        # hang (bench.py leaves the same way)
        os._exit(0)


def _run(mode):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, mode)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        res = sorted((q.get(timeout=240) for _ in range(world)), key=lambda r: r[0])
    finally:
        for p in procs:
            p.join(timeout=30)
            if p.is_alive():
                p.kill()
    return [(r, torch.from_numpy(f), l) for r, f, l in res]


def _single_rank_reference():
    from mamba_b200 import synthetic, train
    model = _small_model().cuda()
    init = torch.cat([p.detach().flatten() for p in model.parameters()]).cpu()
    src, trg, meta = synthetic.batch(4, 96, seed=5)
    tr = train.Trainer(model, lr=1e-3, autocast_dtype=None, world_size=1, batch_size=4, block_len=96)
    losses = [float(tr.step(src.cuda(), trg.cuda(), meta.cuda())) for _ in range(3)]
    torch.cuda.synchronize()
    return init, torch.cat([p.detach().flatten() for p in model.parameters()]).cpu(), losses


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("mode", ["trainer", "ddp"])
def test_two_ranks_equal_one_rank_on_the_concatenated_batch(mode):
    res = _run(mode)
    assert torch.equal(res[0][1], res[1][1]), "ranks hold different parameters after the step"
    init, want, want_losses = _single_rank_reference()
    upd = (want - init).abs().max()
    err = (res[0][1] - want).abs().max()
    print(f"[dist:{mode}] max parameter update {float(upd):.3e}, 2-rank vs 1-rank max abs difference {float(err):.3e}; "
          f"rank-0 losses {res[0][2]} rank-1 {res[1][2]} one-rank {want_losses}")
    diff = (res[0][1] - want).abs()
    q999 = float(torch.quantile(diff[torch.randperm(diff.numel())[:2_000_000]], 0.999))
    print(f"[dist:{mode}] 99.9th percentile of the difference {q999:.3e}")
    assert q999 <= 1e-3 * float(upd) + 1e-8
    assert float(err) <= 2 * 1e-3 * 3 * 1.01
    # the one-rank loss is the mean over the whole batch = the mean of the two ranks' losses
    for a, b, c in zip(res[0][2], res[1][2], want_losses):
        assert abs(0.5 * (a + b) - c) <= 1e-4 * abs(c)
