"""GPU parity tests at module level: MambaBlock / Mamba (both layouts), the training step and the decode
paths against the CPU oracle on identical weights and inputs (fp32 rtol 1e-4; bf16 tolerance stated below)."""
import pytest
import torch

from oracle import simple_mamba as om
from oracle import train_ref
from util import assert_close

pytestmark = pytest.mark.gpu

RTOL32 = 1e-4


def _args(cls, **kw):
    from mamba_b200.configs import common as cc
    base = dict(d_model=64, n_layer=2, vocab_size=cc.vocab_size, d_state=16, expand=2, d_conv=4,
                pad_vocab_size_multiple=1, metadata_vocab_size=cc.metadata_vocab_size)
    base.update(kw)
    return cls(**base)


def _randomise(module, seed):
    """Move every parameter off its init so that A_log, D, biases and norms all matter."""
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


@pytest.mark.parametrize("cfg", [dict(d_model=64, d_state=16), dict(d_model=48, d_state=64), dict(d_model=256, d_state=64)])
@pytest.mark.parametrize("L", [1, 37, 300])
def test_mamba_block_forward_backward_fp32(cfg, L):
    from mamba_b200.models.mamba import MambaBlock, ModelArgs
    torch.manual_seed(0)
    ref = om.MambaBlock(_args(om.ModelArgs, **cfg))
    _randomise(ref, 1)
    blk = MambaBlock(_args(ModelArgs, **cfg)).cuda()
    blk.load_state_dict(ref.state_dict(), strict=True)
    x = torch.randn(2, L, cfg["d_model"])
    dy = torch.randn(2, L, cfg["d_model"])
    xc = x.clone().requires_grad_(True)
    yr = ref(xc)
    yr.backward(dy)
    xg = x.cuda().requires_grad_(True)
    yg = blk(xg)
    yg.backward(dy.cuda())
    assert_close(yg, yr, RTOL32, what="block fwd")
    assert_close(xg.grad, xc.grad, RTOL32, what="block dx")
    gp = dict(blk.named_parameters())
    for name, p in ref.named_parameters():
        # at L = 1 dA_log is exactly 0 in the reference; the kernel leaves one product rounding behind (see
        # test_scan_backward_fp32), hence the absolute floor on that one gradient
        assert_close(gp[name].grad, p.grad, RTOL32, 2e-5, what=f"block d{name}",
                     atol_abs=1e-6 if name == "A_log" else 0.0)


@pytest.mark.parametrize("layout", ["P", "S"])
def test_model_logits_loss_and_grads_fp32(layout):
    """Full model through the reference loss (filtered_logit + CrossEntropy): logits, loss, every gradient."""
    from mamba_b200 import synthetic, train
    from mamba_b200.models.mamba import Mamba, ModelArgs
    torch.manual_seed(0)
    if layout == "P":
        ref = om.Mamba(_args(om.ModelArgs))
        model = Mamba(_args(ModelArgs))
    else:
        # the shipped wrapper: Mamba-2 layers (oracle/mamba2_ref.py restates the published recurrence; parity for this
        # layer is unpinned — mamba_ssm is outside the reference tree)
        from oracle.mamba2_ref import ShippedMambaRef
        model = Mamba(d_model=128, n_layers=2)
        ref = ShippedMambaRef(d_model=128, n_layers=2, d_state=model.params.d_state)
    _randomise(ref, 2)
    model.load_state_dict(ref.state_dict(), strict=True)
    model.cuda()
    src, trg, meta = synthetic.batch(2, 45, seed=5)
    lr = ref(src, meta)
    loss_r = train_ref.loss_fn(src, trg, lr)
    loss_r.backward()
    lg = model(src.cuda(), meta.cuda())
    loss_g = train.loss_fn(src.cuda(), trg.cuda(), lg)
    loss_g.backward()
    assert lg.shape == lr.shape == (2, 45, 17914)
    assert_close(lg, lr, RTOL32, what=f"{layout} logits")
    assert abs(loss_g.item() - loss_r.item()) <= 1e-4 * abs(loss_r.item())
    gp = dict(model.named_parameters())
    gmax = max(float(p.grad.abs().max()) for p in ref.parameters())
    for name, p in ref.named_parameters():
        if float(p.grad.abs().max()) < 1e-6 * gmax:
            # mathematically zero (e.g. output_layer.bias: the sequence-axis log_softmax of F4 removes any
            # per-vocab constant) — both sides hold rounding noise only
            assert float(gp[name].grad.abs().max()) < 1e-5 * gmax, name
            continue
        assert_close(gp[name].grad, p.grad, RTOL32, 2e-5, what=f"{layout} d{name}")


@pytest.mark.parametrize("cfg", [dict(d_model=128, d_state=64), dict(d_model=64, d_state=16, headdim=32)])
def test_mamba2_layer_forward_backward_fp32(cfg):
    """The shipped model's layer (mamba_ssm.Mamba2 as configured at models/mamba/mamba.py:17-23) on the hot path's
    kernels against oracle/mamba2_ref.py (a restatement of the published recurrence: parity UNPINNED, mamba_ssm is
    not in the reference tree): output and every parameter / input gradient, fp32 rtol 1e-4."""
    from mamba_b200.models.mamba.mamba2 import Mamba2
    from oracle.mamba2_ref import Mamba2Ref
    torch.manual_seed(0)
    ref = Mamba2Ref(**cfg)
    _randomise(ref, 3)
    with torch.no_grad():
        ref.dt_bias.copy_(torch.randn_like(ref.dt_bias) - 3)      # softplus(dt + bias) ~ 0.05
        ref.A_log.copy_(torch.log(torch.rand_like(ref.A_log) * 15 + 1))
    lay = Mamba2(**cfg)
    lay.load_state_dict(ref.state_dict(), strict=True)
    lay.cuda()
    x = torch.randn(2, 70, cfg["d_model"], generator=torch.Generator().manual_seed(4))
    xr = x.clone().requires_grad_(True)
    yr = ref(xr)
    g = torch.randn_like(yr)
    yr.backward(g)
    xg = x.cuda().requires_grad_(True)
    yg = lay(xg)
    yg.backward(g.cuda())
    assert_close(yg, yr, RTOL32, what="mamba2 out")
    assert_close(xg.grad, xr.grad, RTOL32, 2e-5, what="mamba2 dx")
    gp = dict(lay.named_parameters())
    for name, p in ref.named_parameters():
        assert_close(gp[name].grad, p.grad, RTOL32, 2e-5, what=f"mamba2 d{name}")


def test_layout_s_prefill_and_step_match_full_forward():
    """Recurrent decode of the shipped (Mamba-2) layout: prefill + step logits equal the full forward's."""
    from mamba_b200 import synthetic
    from mamba_b200.models.mamba import Mamba
    torch.manual_seed(0)
    model = Mamba(d_model=128, n_layers=2).cuda().eval()
    src, _, meta = synthetic.batch(3, 40, seed=6)
    with torch.no_grad():
        full = model(src.cuda(), meta.cuda())
        cache = model.allocate_inference_cache(3)
        pre = model.prefill(src[:, :25].cuda(), meta.cuda(), cache)
        assert_close(pre, full[:, :25], RTOL32, 2e-5, what="layout S prefill logits")
        for t in range(25, 40):
            lg = model.step(src[:, t].cuda(), cache)
            assert_close(lg, full[:, t], RTOL32, 2e-5, what=f"layout S step logits t={t}")


def test_model_bf16_autocast_tolerance():
    """bf16 mixer / fp32 residual stream under autocast vs the fp32 oracle: logits within 3e-2 of max|logit|
    (stated bf16 tolerance: two layers of bf16 GEMMs with 2^-9 input rounding each)."""
    from mamba_b200 import synthetic
    from mamba_b200.models.mamba import Mamba, ModelArgs
    torch.manual_seed(0)
    ref = om.Mamba(_args(om.ModelArgs))
    _randomise(ref, 3)
    model = Mamba(_args(ModelArgs))
    model.load_state_dict(ref.state_dict())
    model.cuda()
    src, trg, meta = synthetic.batch(2, 64, seed=6)
    with torch.no_grad():
        lr = ref(src, meta)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            lg = model(src.cuda(), meta.cuda())
    assert lg.dtype == torch.bfloat16
    err = (lg.float().cpu() - lr).abs().max().item()
    assert err <= 3e-2 * lr.abs().max().item(), err


def test_state_dict_roundtrip_on_gpu(tmp_path):
    from mamba_b200.models.mamba import Mamba, ModelArgs
    m1 = Mamba(_args(ModelArgs)).cuda()
    torch.save(m1.state_dict(), tmp_path / "m.pth")       # train.py:77
    m2 = Mamba(_args(ModelArgs)).cuda()
    m2.load_state_dict(torch.load(tmp_path / "m.pth"), strict=True)  # train.py:66
    for (k1, v1), (k2, v2) in zip(m1.state_dict().items(), m2.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)


def test_prefill_and_step_match_full_forward():
    """Row A9: prefill == forward, and stepping token by token reproduces the forward logits position by position."""
    from mamba_b200 import synthetic
    from mamba_b200.models.mamba import Mamba, ModelArgs
    torch.manual_seed(0)
    ref = om.Mamba(_args(om.ModelArgs))
    _randomise(ref, 4)
    model = Mamba(_args(ModelArgs))
    model.load_state_dict(ref.state_dict())
    model.cuda().eval()
    src, _, meta = synthetic.batch(3, 40, seed=7)
    with torch.no_grad():
        full_ref = ref(src, meta)
        full = model(src.cuda(), meta.cuda())
        cache = model.allocate_inference_cache(3)
        pre = model.prefill(src[:, :25].cuda(), meta.cuda(), cache)
        assert_close(pre, full_ref[:, :25], RTOL32, what="prefill logits")
        assert torch.equal(pre, full[:, :25])
        for t in range(25, 40):
            lg = model.step(src[:, t].cuda(), cache)
            assert_close(lg, full_ref[:, t], RTOL32, 2e-5, what=f"step logits t={t}")


def test_greedy_decode_tokens_match_oracle():
    """Greedy decode (scripts/generate_midi_many.py:13-56): literal loop on the kernels, recurrent decoder (with and
    without the CUDA graph) and the oracle loop produce the same token sequence — strict equality, no exemptions."""
    from mamba_b200 import generate, synthetic
    from mamba_b200.models.mamba import Mamba, ModelArgs
    torch.manual_seed(0)
    ref = om.Mamba(_args(om.ModelArgs)).eval()
    _randomise(ref, 5)
    model = Mamba(_args(ModelArgs))
    model.load_state_dict(ref.state_dict())
    model.cuda().eval()
    src, _, meta = synthetic.batch(2, 120, seed=8)
    n_new = 24
    want = [train_ref.generate_greedy(ref, 4096, src[i:i + 1].clone(), meta[i:i + 1], n_new) for i in range(2)]
    want = torch.tensor(want)
    lit = generate.generate_literal(model, 4096, src.cuda(), meta.cuda(), n_new).cpu()
    rec = generate.generate_recurrent(model, src.cuda(), meta.cuda(), n_new, use_graph=True).cpu()
    rec_nograph = generate.generate_recurrent(model, src.cuda(), meta.cuda(), n_new, use_graph=False).cpu()
    assert torch.equal(rec, rec_nograph)
    assert torch.equal(lit, want), (lit[:, 120:], want[:, 120:])
    assert torch.equal(rec, want), (rec[:, 120:], want[:, 120:])


def test_sampling_decode_tokens_match_oracle_given_the_same_uniforms():
    """scripts/generate.py:14-95 (look-back window bounded by summed time shifts, k by token class, penalties, top-k,
    one draw): the recurrent decoder with the on-device sampler (csrc/sample.cu) against the oracle's restatement of
    the loop on the CPU oracle model, both fed the same uniforms.  Parity = identical tokens."""
    from mamba_b200 import generate, synthetic
    from mamba_b200.models.mamba import Mamba, ModelArgs
    torch.manual_seed(0)
    ref = om.Mamba(_args(om.ModelArgs)).eval()
    _randomise(ref, 6)
    model = Mamba(_args(ModelArgs))
    model.load_state_dict(ref.state_dict())
    model.cuda().eval()
    src, _, meta = synthetic.batch(3, 90, seed=9)
    n_new = 24
    U = torch.rand(n_new + 1, 3, 2, generator=torch.Generator().manual_seed(77))
    want = torch.tensor(train_ref.generate_sampling(ref, 4096, src.clone(), meta, n_new, U))
    for use_graph in (False, True):
        got = generate.generate_recurrent(model, src.cuda(), meta.cuda(), n_new, use_graph=use_graph, mode="sample",
                                          uniforms=U.cuda()).cpu()
        assert torch.equal(got, want), (use_graph, got[:, 90:], want[:, 90:])
    # the drop-in entry point (scripts/generate.py signature) runs and is reproducible under a seed
    a = generate.generate(model, 4096, src, meta, num_tokens=8, device="cuda", seed=3)
    b = generate.generate(model, 4096, src, meta, num_tokens=8, device="cuda", seed=3)
    assert a == b and len(a) == 3 and len(a[0]) == 98


def test_fused_decode_step_equals_unfused_step():
    """Mamba.step's fused path (norm + conv folded into in_proj, final norm into the head; 4 launches per layer)
    against the one-kernel-per-op path on the same weights and states, fp32 and bf16."""
    from mamba_b200 import synthetic
    from mamba_b200.models.mamba import Mamba, ModelArgs
    for dt, rtol, floor in ((torch.float32, 1e-4, 2e-5), (torch.bfloat16, 2e-2, 2e-2)):
        torch.manual_seed(0)
        model = Mamba(_args(ModelArgs)).cuda().eval().to(dt)
        src, _, meta = synthetic.batch(3, 40, seed=12)
        c1, c2 = model.allocate_inference_cache(3), model.allocate_inference_cache(3)
        c2.fused = False
        with torch.no_grad():
            model.prefill(src[:, :20].cuda(), meta.cuda(), c1)
            model.prefill(src[:, :20].cuda(), meta.cuda(), c2)
            for t in range(20, 40):
                a = model.step(src[:, t].cuda(), c1)
                b = model.step(src[:, t].cuda(), c2)
                assert_close(a, b, rtol, floor, what=f"fused step logits {dt} t={t}")
        for (cs1, hs1), (cs2, hs2) in zip(c1, c2):
            assert_close(cs1, cs2, rtol, floor, what="conv state")
            assert_close(hs1, hs2, rtol, floor, what="ssm state")


def test_persistent_decode_kernel_equals_step_by_step_path():
    """mamba_decode_token (the whole model as ONE cooperative launch per token, csrc/decode.cu) against the
    kernel-per-op decode path on the same weights: logits of every step within fp32 tolerance, identical greedy
    tokens, identical final states; then the bf16-weight variant against the fp32 one at bf16 tolerance."""
    from mamba_b200 import generate, synthetic
    from mamba_b200.configs import common as cc
    from mamba_b200.models.mamba import Mamba, ModelArgs
    torch.manual_seed(0)
    model = Mamba(ModelArgs(d_model=128, n_layer=3, vocab_size=cc.vocab_size, d_state=16, expand=2, d_conv=4,
                            pad_vocab_size_multiple=1, metadata_vocab_size=cc.metadata_vocab_size))
    _randomise(model, 7)
    model.cuda().eval()
    src, _, meta = synthetic.batch(5, 60, seed=13)
    src, meta = src.cuda(), meta.cuda()
    decs = {}
    for name, kw in (("steps", dict(persistent=False)), ("one", dict(persistent=True)),
                     ("one_graph", dict(persistent=True, use_graph=True)), ("one_bf16", dict(persistent=True, weight_dtype=torch.bfloat16))):
        kw.setdefault("use_graph", False)
        d = generate.RecurrentDecoder(model, 5, max_new_tokens=64, **kw)
        d.prefill(src, meta)
        decs[name] = d
    assert decs["one"].plan is not None and decs["steps"].plan is None
    for t in range(40):
        for d in decs.values():
            d.step()
        ref = decs["steps"]
        assert_close(decs["one"].logits, ref.logits, 1e-4, 2e-5, what=f"persistent decode logits, step {t}")
        assert torch.equal(decs["one_graph"].logits, decs["one"].logits)
        assert torch.equal(decs["one"].nxt, ref.nxt), t
        if t < 3:   # before the sequences can part ways
            assert_close(decs["one_bf16"].logits, ref.logits, 2e-2, 2e-2, what=f"bf16-weight decode logits, step {t}")
    for (cs1, hs1), (cs2, hs2) in zip(decs["one"].cache, decs["steps"].cache):
        assert_close(cs1, cs2, 1e-4, 2e-5, what="conv state")
        assert_close(hs1, hs2, 1e-4, 2e-5, what="ssm state")
    assert torch.equal(decs["one"].tokens(), decs["steps"].tokens())


@pytest.mark.parametrize("nseq", [1, 2, 3, 4, 7, 10, 16])
def test_persistent_decode_kernel_every_batch_size(nseq):
    """Batch sizes 1..16 (the kernel is instantiated per even batch; odd ones run with one padded sequence; sequence-
    sharded generation leaves a rank with 1, 2, 3 or 5 of the 10 sequences): persistent kernel against the
    kernel-per-op path, logits within fp32 tolerance at every step and identical greedy tokens."""
    from mamba_b200 import generate, synthetic
    from mamba_b200.configs import common as cc
    from mamba_b200.models.mamba import Mamba, ModelArgs
    torch.manual_seed(1)
    model = Mamba(ModelArgs(d_model=128, n_layer=2, vocab_size=cc.vocab_size, d_state=16, expand=2, d_conv=4,
                            pad_vocab_size_multiple=1, metadata_vocab_size=cc.metadata_vocab_size))
    _randomise(model, 11)
    model.cuda().eval()
    src, _, meta = synthetic.batch(nseq, 40, seed=17 + nseq)
    src, meta = src.cuda(), meta.cuda()
    a = generate.RecurrentDecoder(model, nseq, max_new_tokens=32, persistent=False, use_graph=False)
    b = generate.RecurrentDecoder(model, nseq, max_new_tokens=32, persistent=True, use_graph=False)
    a.prefill(src, meta), b.prefill(src, meta)
    assert b.plan is not None
    for t in range(12):
        a.step(), b.step()
        assert_close(b.logits, a.logits, 1e-4, 2e-5, what=f"persistent decode logits, batch {nseq}, step {t}")
        assert torch.equal(a.nxt, b.nxt), (nseq, t)
    for (cs1, hs1), (cs2, hs2) in zip(b.cache, a.cache):
        assert_close(cs1, cs2, 1e-4, 2e-5, what="conv state")
        assert_close(hs1, hs2, 1e-4, 2e-5, what="ssm state")


def test_trainer_graph_step_equals_eager_step():
    """The CUDA-graphed step (Trainer) and the python-launched reference-shaped step produce the same losses."""
    from mamba_b200 import synthetic, train
    from mamba_b200.models.mamba import Mamba, ModelArgs
    torch.manual_seed(0)
    a = Mamba(_args(ModelArgs)).cuda()
    b = Mamba(_args(ModelArgs)).cuda()
    b.load_state_dict(a.state_dict())
    opt = torch.optim.Adam(a.parameters(), lr=1e-3)
    tr = train.Trainer(b, lr=1e-3, autocast_dtype=None, batch_size=2, block_len=48, use_graph=True)
    for i in range(4):
        src, trg, meta = (t.cuda() for t in synthetic.batch(2, 48, seed=20 + i))
        la = train.train_step(a, opt, src, trg, meta)
        lb = tr.step(src, trg, meta)
        assert abs(la.item() - lb.item()) <= 2e-4 * abs(la.item()), (i, la.item(), lb.item())
    for (n1, p1), (n2, p2) in zip(a.named_parameters(), b.named_parameters()):
        assert_close(p2, p1, 1e-3, 1e-4, what=f"param {n1} after 4 steps")


@pytest.mark.parametrize("use_graph", [False, True])
@pytest.mark.parametrize("stages", [2, 4])
def test_staged_backward_with_side_stream_adam_equals_plain_step(use_graph, stages):
    """One GPU: the staged step (backward cut into layer groups, each group's Adam update on a side stream under
    the backward of the groups below) trains bit-identically to backward-then-Adam."""
    from mamba_b200 import synthetic, train
    from mamba_b200.models.mamba import Mamba, ModelArgs
    from mamba_b200.models.mamba.mamba import AsyncWgrad
    torch.manual_seed(0)
    a = Mamba(_args(ModelArgs, d_model=64, d_state=64, n_layer=4)).cuda()
    b = Mamba(_args(ModelArgs, d_model=64, d_state=64, n_layer=4)).cuda()
    b.load_state_dict(a.state_dict())
    try:
        ta = train.Trainer(a, lr=1e-3, batch_size=2, block_len=40, use_graph=use_graph, stages=1)
        tb = train.Trainer(b, lr=1e-3, batch_size=2, block_len=40, use_graph=use_graph, stages=stages)
        assert ta._stage_groups is None and len(tb._stage_groups) == stages and len(tb._stage_optimizers) == stages + 1
        for i in range(3):
            batch = [t.cuda() for t in synthetic.batch(2, 40, seed=40 + i)]
            la, lb = ta.step(*batch).item(), tb.step(*batch).item()
            assert la == lb, (i, la, lb)
    finally:
        AsyncWgrad.disable()
    for (n1, p1), (n2, p2) in zip(a.named_parameters(), b.named_parameters()):
        assert torch.equal(p1, p2), f"param {n1} differs between the staged and the plain step"


@pytest.mark.parametrize("use_graph", [False, True])
def test_side_stream_weight_gradients_equal_inline_ones(use_graph):
    """bf16-autocast training steps with the weight-gradient GEMMs on the side stream (AsyncWgrad, written straight
    into param.grad) and with them inline: same losses, same parameters after 3 Adam steps (same GEMMs, same
    inputs -> bit-identical)."""
    from mamba_b200 import synthetic, train
    from mamba_b200.models.mamba import Mamba, ModelArgs
    from mamba_b200.models.mamba.mamba import AsyncWgrad
    torch.manual_seed(0)
    a = Mamba(_args(ModelArgs, d_model=128, d_state=64)).cuda()
    b = Mamba(_args(ModelArgs, d_model=128, d_state=64)).cuda()
    b.load_state_dict(a.state_dict())
    try:
        ta = train.Trainer(a, lr=1e-3, batch_size=2, block_len=48, use_graph=use_graph, async_wgrad=False)
        assert not ta.async_wgrad and ta._wgrad_stream is None
        la = [ta.step(*(t.cuda() for t in synthetic.batch(2, 48, seed=30 + i))).item() for i in range(3)]
        tb = train.Trainer(b, lr=1e-3, batch_size=2, block_len=48, use_graph=use_graph, async_wgrad=True)
        assert tb.async_wgrad and tb._wgrad_stream is not None and len(tb._wgrad_params) == 4 * 2
        # the side-stream path is scoped to the Trainer's own backward: nothing is armed process-wide
        assert AsyncWgrad.stream is None and not AsyncWgrad.owned
        lb = [tb.step(*(t.cuda() for t in synthetic.batch(2, 48, seed=30 + i))).item() for i in range(3)]
    finally:
        AsyncWgrad.disable()
    assert la == lb, (la, lb)
    for (n1, p1), (n2, p2) in zip(a.named_parameters(), b.named_parameters()):
        assert torch.equal(p1, p2), f"param {n1} differs between inline and side-stream weight gradients"


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("padded", [False, True])
def test_fused_loss_matches_oracle_loss(dtype, padded):
    """ops.filtered_ce_fn (csrc/loss.cu) against the oracle's train.py:133-138 + :161-165, value and gradient; also
    through a row-padded logits view (the LM head's layout)."""
    from mamba_b200 import synthetic, train
    from mamba_b200.configs import common as cc
    V = cc.vocab_size
    torch.manual_seed(0)
    src, trg, _ = synthetic.batch(2, 150, seed=9)
    base = torch.randn(2, 150, V + 6) * 2.0
    x = base.to(dtype)
    xr = x[..., :V].float().clone().requires_grad_(True)
    loss_r = train_ref.loss_fn(src, trg, xr)
    (loss_r * 1.7).backward()
    if padded:
        xg_base = x.cuda().requires_grad_(True)
        xg = xg_base[..., :V]
    else:
        xg_base = x[..., :V].contiguous().cuda().requires_grad_(True)
        xg = xg_base
    loss_g = train.loss_fn(src.cuda(), trg.cuda(), xg)
    (loss_g * 1.7).backward()
    assert abs(loss_g.item() - loss_r.item()) <= 1e-4 * abs(loss_r.item()), (loss_g.item(), loss_r.item())
    got = xg_base.grad[..., :V]
    if dtype == torch.float32:
        assert_close(got, xr.grad, 1e-4, 1e-5, what="loss dlogits fp32")
    else:
        assert_close(got, xr.grad, 2e-2, 2e-2, what="loss dlogits bf16")
    # torch-op spelling of the same loss on the GPU (the drop-in mirror of the reference functions)
    loss_t = train.loss_fn_torch(src.cuda(), trg.cuda(), xg.detach().float())
    assert abs(loss_t.item() - loss_r.item()) <= 1e-4 * abs(loss_r.item())


def test_block_matches_reference_golden_fixture():
    """CUDA MambaBlock against outputs of the reference's own bytecode (tests/golden/block_small.pt)."""
    from pathlib import Path
    from mamba_b200.models.mamba import MambaBlock, ModelArgs
    f = torch.load(Path(__file__).resolve().parent / "golden" / "block_small.pt")
    blk = MambaBlock(ModelArgs(d_model=32, n_layer=1, vocab_size=64, d_state=16))
    blk.load_state_dict(f["state"], strict=True)
    blk.cuda()
    x = f["x"].cuda().requires_grad_(True)
    y = blk(x)
    y.backward(f["dy"].cuda())
    assert_close(y, f["y"], RTOL32, what="CUDA block fwd vs reference golden")
    assert_close(x.grad, f["dx"], RTOL32, what="CUDA block dx vs reference golden")
    for k, p in blk.named_parameters():
        assert_close(p.grad, f["grads"][k], RTOL32, 2e-5, what=f"CUDA block d{k} vs reference golden")


def test_model_and_loss_match_reference_golden_fixture():
    """CUDA Mamba (Layout P) + fused loss against the reference's model bytecode and train.py loss source."""
    from pathlib import Path
    from mamba_b200 import synthetic, train
    from mamba_b200.configs import common as cc
    from mamba_b200.models.mamba import Mamba, ModelArgs
    f = torch.load(Path(__file__).resolve().parent / "golden" / "model_small.pt")
    torch.manual_seed(f["init_seed"])
    ref = om.Mamba(om.ModelArgs(vocab_size=cc.vocab_size, pad_vocab_size_multiple=1, **f["params"]))
    _golden_randomise(ref, f["rand_seed"])
    model = Mamba(ModelArgs(vocab_size=cc.vocab_size, pad_vocab_size_multiple=1, **f["params"]))
    model.load_state_dict(ref.state_dict(), strict=True)
    model.cuda()
    src, trg, meta = synthetic.batch(2, 20, seed=f["batch_seed"])
    logits = model(src.cuda(), meta.cuda())
    loss = train.loss_fn(src.cuda(), trg.cuda(), logits)
    loss.backward()
    assert_close(logits[:, :, ::97], f["logits_sample"], RTOL32, what="CUDA logits vs reference golden")
    assert abs(loss.item() - f["loss"].item()) <= 1e-4 * abs(f["loss"].item())
    grads = dict(model.named_parameters())
    gmax = max(float(g.abs().max()) for g in f["grads"].values())
    for k, g in f["grads"].items():
        assert_close(grads[k].grad, g, RTOL32, 2e-5, what=f"CUDA model d{k} vs reference golden", atol_abs=1e-7 * gmax)
    assert_close(model.embedding.weight.grad[f["emb_rows"].cuda()], f["emb_grad_rows"], RTOL32, 2e-5,
                 what="CUDA embedding grad rows vs reference golden")


def _golden_randomise(module, seed):  # identical to tests/golden/make_golden.py
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


def test_trainer_resume_keeps_loaded_adam_state_through_capture():
    """Trainer.capture() (lazy, on the first step) must not wipe optimizer state loaded through
    Trainer.optimizers[i].load_state_dict() — the documented resume path.  Train 2 steps, save parameters and Adam
    state, build a fresh graph-captured Trainer from them, step once more: same result as continuing."""
    from mamba_b200 import synthetic, train
    from mamba_b200.models.mamba import Mamba, ModelArgs
    torch.manual_seed(0)
    a = Mamba(_args(ModelArgs, d_model=64, d_state=64, n_layer=4)).cuda()
    ta = train.Trainer(a, lr=1e-3, batch_size=2, block_len=40, use_graph=True, stages=2)
    batches = [[t.cuda() for t in synthetic.batch(2, 40, seed=60 + i)] for i in range(3)]
    for i in range(2):
        ta.step(*batches[i])
    torch.cuda.synchronize()
    model_sd = {k: v.clone() for k, v in a.state_dict().items()}
    opt_sd = [o.state_dict() for o in ta.optimizers]
    import copy
    opt_sd = copy.deepcopy(opt_sd)
    la = ta.step(*batches[2]).item()

    b = Mamba(_args(ModelArgs, d_model=64, d_state=64, n_layer=4)).cuda()
    b.load_state_dict(model_sd)
    tb = train.Trainer(b, lr=1e-3, batch_size=2, block_len=40, use_graph=True, stages=2)
    for o, sd in zip(tb.optimizers, opt_sd):
        o.load_state_dict(sd)
    lb = tb.step(*batches[2]).item()
    assert la == lb, (la, lb)
    for (n1, p1), (n2, p2) in zip(a.named_parameters(), b.named_parameters()):
        assert_close(p2, p1, 1e-6, 1e-7, what=f"param {n1} after the resumed step")
    # the resumed optimizer's step counter went from 2 to 3, not from 0 to 1
    steps = {float(st["step"]) for o in tb.optimizers for st in o.state.values()}
    assert steps == {3.0}, steps


def test_side_stream_wgrad_is_scoped_to_its_trainer():
    """with a Trainer alive (async weight gradients on), a plain train_step on ANOTHER model — and
    gradient accumulation over two micro-batches — must take the ordinary autograd path: gradients accumulate."""
    from mamba_b200 import synthetic, train
    from mamba_b200.models.mamba import Mamba, ModelArgs
    from mamba_b200.models.mamba.mamba import AsyncWgrad
    torch.manual_seed(0)
    owner = Mamba(_args(ModelArgs, d_model=64, d_state=64)).cuda()
    tr = train.Trainer(owner, lr=1e-3, batch_size=2, block_len=32, use_graph=False, async_wgrad=True)
    tr.step(*(t.cuda() for t in synthetic.batch(2, 32, seed=70)))
    assert tr.async_wgrad and AsyncWgrad.stream is None
    other = Mamba(_args(ModelArgs, d_model=64, d_state=64)).cuda()
    src, trg, meta = (t.cuda() for t in synthetic.batch(2, 32, seed=71))

    def grads_after(n_micro):
        other.zero_grad(set_to_none=True)
        for _ in range(n_micro):
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out = other(src, meta)
            train.loss_fn(src, trg, out).backward()
        torch.cuda.synchronize()
        return {n: p.grad.clone() for n, p in other.named_parameters()}

    g1, g2 = grads_after(1), grads_after(2)
    for n in g1:
        assert_close(g2[n], 2 * g1[n], 1e-5, 1e-6, what=f"accumulated gradient of {n}")
