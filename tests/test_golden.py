"""The oracle restatement (oracle/simple_mamba.py, oracle/train_ref.py) against the golden fixtures produced by the
REFERENCE ITSELF (tests/golden/make_golden.py executes the reference's own bytecode / source), and — when
/root/reference is mounted — against a live execution of that bytecode, bit for bit."""
from pathlib import Path

import pytest
import torch

from oracle import reference_exec as rx
from oracle import simple_mamba as om
from oracle import train_ref
from util import assert_close

GOLD = Path(__file__).resolve().parent / "golden"
TIGHT = dict(rtol=1e-6, atol_frac=1e-6)  # same arithmetic, same order; slack only for a different host CPU's BLAS


def randomise(module, seed):  # identical to tests/golden/make_golden.py
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


@pytest.mark.parametrize("impl", ["literal", "unbind"])
def test_selective_scan_matches_reference_fixture(impl):
    f = torch.load(GOLD / "scan_small.pt")
    y = om.selective_scan(f["u"], f["delta"], f["A"], f["B"], f["C"], f["D"], impl=impl)
    assert_close(y, f["y"], what=f"oracle scan ({impl}) vs reference", **TIGHT)


def test_block_matches_reference_fixture():
    f = torch.load(GOLD / "block_small.pt")
    blk = om.MambaBlock(om.ModelArgs(d_model=32, n_layer=1, vocab_size=64, d_state=16))
    blk.load_state_dict(f["state"], strict=True)
    x = f["x"].clone().requires_grad_(True)
    y = blk(x)
    y.backward(f["dy"])
    assert_close(y, f["y"], what="oracle block fwd vs reference", **TIGHT)
    assert_close(x.grad, f["dx"], what="oracle block dx vs reference", **TIGHT)
    for k, p in blk.named_parameters():
        assert_close(p.grad, f["grads"][k], what=f"oracle block d{k} vs reference", **TIGHT)


def test_model_and_loss_match_reference_fixture():
    from mamba_b200 import synthetic
    f = torch.load(GOLD / "model_small.pt")
    V, _ = train_ref.vocab_layout()
    torch.manual_seed(f["init_seed"])
    model = om.Mamba(om.ModelArgs(vocab_size=V, pad_vocab_size_multiple=1, **f["params"]))
    randomise(model, f["rand_seed"])
    src, trg, meta = synthetic.batch(2, 20, seed=f["batch_seed"])
    logits = model(src, meta)
    loss = train_ref.loss_fn(src, trg, logits)
    loss.backward()
    assert torch.equal(train_ref.make_distributions(), f["distributions"])
    assert_close(logits[:, :, ::97], f["logits_sample"], what="oracle logits vs reference", **TIGHT)
    assert abs(float(logits.double().sum()) - float(f["logits_sum"])) <= 1e-6 * abs(float(f["logits_sum"])) + 1e-4
    assert abs(loss.item() - f["loss"].item()) <= 1e-6 * abs(f["loss"].item())
    grads = dict(model.named_parameters())
    for k, g in f["grads"].items():
        assert_close(grads[k].grad, g, what=f"oracle model d{k} vs reference", **TIGHT)
    assert_close(model.embedding.weight.grad[f["emb_rows"]], f["emb_grad_rows"], what="oracle embedding grad rows",
                 **TIGHT)


@pytest.mark.skipif(not rx.available(), reason="/root/reference is not mounted (GPU box)")
def test_restatement_is_bit_identical_to_live_reference_bytecode():
    """Executes the reference's 3.11 bytecode here and now: same init under the same seed, same forward, same
    gradients, bit for bit, for the block and for the full model."""
    ref = rx.load_reference()
    p = rx.make_params(ref, d_model=32, n_layer=2, vocab_size=100, d_state=8)
    assert (p.d_inner, p.dt_rank, p.vocab_size) == (64, 2, 104)   # ModelArgs.__post_init__ @L46-54 executed
    torch.manual_seed(1)
    rb = ref.MambaBlock(p)
    torch.manual_seed(1)
    mb = om.MambaBlock(om.ModelArgs(d_model=32, n_layer=2, vocab_size=100, d_state=8), scan_impl="literal")
    assert list(rb.state_dict()) == list(mb.state_dict())
    assert all(torch.equal(a, b) for a, b in zip(rb.state_dict().values(), mb.state_dict().values()))
    x = torch.randn(2, 20, 32)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ya, yb = rb(xa), mb(xb)
    assert torch.equal(ya, yb)
    ya.square().sum().backward()
    yb.square().sum().backward()
    assert torch.equal(xa.grad, xb.grad)
    assert all(torch.equal(a.grad, b.grad) for a, b in zip(rb.parameters(), mb.parameters()))
    rm = ref.Mamba(p)
    mm = om.Mamba(om.ModelArgs(d_model=32, n_layer=2, vocab_size=100, d_state=8), "unbind")
    assert list(rm.state_dict()) == list(mm.state_dict())
    mm.load_state_dict(rm.state_dict())
    tok, meta = torch.randint(0, 100, (2, 15)), torch.randint(0, 568, (2, 6))
    assert torch.equal(rm(tok, meta), mm(tok, meta))
    rn, mn = ref.RMSNorm(32), om.RMSNorm(32)
    assert rn.eps == mn.eps == 1e-5 and torch.equal(rn(x), mn(x))
