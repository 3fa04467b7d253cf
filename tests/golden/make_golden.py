#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ from the REFERENCE ITSELF (run in the build container only).

The reference's pure-PyTorch Mamba exists only as CPython-3.11 bytecode; `oracle/reference_exec.py` executes that
bytecode (selective_scan, ssm, MambaBlock.forward/__init__, RMSNorm, ResidualBlock, Mamba.forward) against real
torch.  The loss functions are taken from the reference's `train.py` source by extracting the three function
definitions with `ast` and executing them with a stand-in `cc` (their module has import-time side effects: it opens
/scratch/... and calls .to("cuda")).  Nothing of the reference is copied into the repository: only inputs (or
their seeds) and the reference's outputs are stored.

    python tests/golden/make_golden.py          # rewrites tests/golden/*.pt

/root/reference does not exist on the GPU box; the tests only read the committed .pt files.
"""
import ast
import sys
from pathlib import Path
from types import SimpleNamespace

import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import reference_exec as rx  # noqa: E402

OUT = Path(__file__).resolve().parent
REF_TRAIN = Path("/root/reference/train.py")


def randomise(module, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in module.named_parameters():
            if name.endswith("A_log"):
                p.add_(0.2 * torch.randn(p.shape, generator=g))
            elif name.endswith("dt_proj.bias"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.5 - 3.0)
            elif p.dim() == 1:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
            else:
                p.add_(0.02 * torch.randn(p.shape, generator=g))


def reference_loss_functions():
    """train.py:79-138 executed from the reference's source text with stand-in globals."""
    tree = ast.parse(REF_TRAIN.read_text())
    wanted = {"make_distributions", "pick_distributions_by_prev_token", "filtered_logit"}
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in wanted]
    assert {n.name for n in body} == wanted
    disc = SimpleNamespace(pitch=128, dyn=128, length=512, time=512, channel=129, tempo=250)  # configs/common/config.yaml:1-7
    sizes = [disc.pitch * disc.channel, disc.dyn, disc.length, disc.time, disc.tempo]
    start, off = {}, 0
    for k, s in zip(("pitch", "dyn", "length", "time", "tempo"), sizes):
        start[k] = off
        off += s
    cc = SimpleNamespace(vocab_size=sum(sizes), start_idx=start,
                         config=SimpleNamespace(values=SimpleNamespace(block_len=2048, device="cpu"), discretization=disc))
    g = dict(torch=torch, F=F, cc=cc, length_tensor=torch.linspace(1, 3, steps=disc.length - 1))  # train.py:20
    exec(compile(ast.Module(body=body, type_ignores=[]), str(REF_TRAIN), "exec"), g)
    return SimpleNamespace(**{k: g[k] for k in wanted}, cc=cc)


def main():
    ref = rx.load_reference()
    torch.set_num_threads(1)

    # (a) selective_scan alone: MambaBlock.selective_scan @L283-333
    p = rx.make_params(ref, d_model=24, n_layer=1, vocab_size=64, d_state=16)
    blk = ref.MambaBlock(p)
    g = torch.Generator().manual_seed(11)
    Bsz, L, D, N = 2, 45, p.d_inner, p.d_state
    u = torch.randn(Bsz, L, D, generator=g)
    delta = F.softplus(torch.randn(Bsz, L, D, generator=g) - 3)
    A = -torch.exp(torch.randn(D, N, generator=g) * 0.3) * torch.arange(1, N + 1)
    Bm, Cm = torch.randn(Bsz, L, N, generator=g), torch.randn(Bsz, L, N, generator=g)
    Dv = torch.randn(D, generator=g)
    y = blk.selective_scan(u, delta, A, Bm, Cm, Dv)
    torch.save(dict(u=u, delta=delta, A=A, B=Bm, C=Cm, D=Dv, y=y), OUT / "scan_small.pt")

    # (b) one MambaBlock, forward + every gradient: @L185-245 (+ ssm, selective_scan)
    torch.manual_seed(3)
    p = rx.make_params(ref, d_model=32, n_layer=1, vocab_size=64, d_state=16)
    blk = ref.MambaBlock(p)
    randomise(blk, 5)
    x = torch.randn(2, 37, 32, generator=g).requires_grad_(True)
    dy = torch.randn(2, 37, 32, generator=g)
    out = blk(x)
    out.backward(dy)
    torch.save(dict(params=dict(d_model=32, d_state=16), state=blk.state_dict(), x=x.detach(), dy=dy, y=out.detach(),
                    dx=x.grad, grads={k: v.grad for k, v in blk.named_parameters()}), OUT / "block_small.pt")

    # (c) the whole model (Layout P) + the reference loss: Mamba.forward @L74-96, train.py:133-138, :161-165
    lossf = reference_loss_functions()
    V = lossf.cc.vocab_size
    torch.manual_seed(7)
    p = rx.make_params(ref, d_model=16, n_layer=2, vocab_size=V, d_state=8, pad_vocab_size_multiple=1)
    model = ref.Mamba(p)
    randomise(model, 9)
    from mamba_b200 import synthetic
    src, trg, meta = synthetic.batch(2, 20, seed=4)
    logits = model(src, meta)
    filtered = lossf.filtered_logit(src, logits).reshape(-1, V)
    loss = torch.nn.CrossEntropyLoss()(filtered, trg.view(-1))  # train.py:165
    loss.backward()
    keep = [n for n, _ in model.named_parameters() if "embedding" not in n]   # the two [V, d] tables stay out (size)
    grads = {n: q.grad for n, q in model.named_parameters() if n in keep}
    emb_rows = torch.unique(src)
    torch.save(dict(params=dict(d_model=16, n_layer=2, d_state=8), init_seed=7, rand_seed=9, batch_seed=4,
                    logits_sample=logits.detach()[:, :, ::97].clone(), logits_sum=logits.detach().double().sum(),
                    loss=loss.detach(), grads=grads, emb_rows=emb_rows,
                    emb_grad_rows=model.embedding.weight.grad[emb_rows].clone(),
                    distributions=lossf.make_distributions()), OUT / "model_small.pt")
    for f in sorted(OUT.glob("*.pt")):
        print(f.name, f.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
