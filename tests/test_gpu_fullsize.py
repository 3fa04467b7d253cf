"""GPU parity tests AT THE BASELINE SHAPES (VERDICT r01 "next round" item 1).

The checker is the oracle's own torch code (oracle/simple_mamba.py, the restatement proven bit-identical to the
reference's bytecode by tests/test_golden.py) evaluated in fp32 on the GPU: at B=2, L=2054, D=2048, N=64 its
[B, L, D, N] temporaries are 2 x 2.15 GB plus the autograd tape of the python loop — minutes on the host cores,
seconds on the device.  TF32 is switched off for the checker.  The product path under test is the C-ABI library.

Shapes (BASELINE.json configs):
  (a) training shape, bf16 I/O  : B=2, L=2054, D=2048, N=64 — the fused backward the bench line is credited on
  (b) training shape, fp32 I/O  : same, through the lane<->channel backward
  (c) config 5, fp32            : B=2, L=8192, D=2048, N=16, forward and backward
  (d) 10-layer d_model=1024 model, one bf16-autocast training step: loss and every parameter gradient vs the fp32
      oracle (per-layer activation checkpointing in the oracle, as BASELINE.md section 2 plans for the CPU arm)
  (e) greedy decode, 10 sequences x 2000 new tokens on the default model: recurrent decoder vs the literal
      full-re-forward loop of scripts/generate_midi_many.py:13-56

Tolerances (north_star): fp32 rtol 1e-4 with an absolute floor of 1e-5 * max|ref| (2e-5 for gradients that are sums
over all B*L timesteps); bf16 I/O rtol 2e-2 with floor 2e-2 * max|ref|; bf16-autocast model: stated per assertion.
"""
import pytest
import torch
import torch.nn.functional as F

from oracle import simple_mamba as om
from oracle import train_ref
from util import assert_close, scan_inputs

pytestmark = pytest.mark.gpu

RTOL32 = 1e-4
RTOL16, FLOOR16 = 2e-2, 2e-2


@pytest.fixture(autouse=True)
def _no_tf32():
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    torch.cuda.empty_cache()


def _oracle_on_gpu(t, dout):
    """Oracle evaluation (fp32, on the device) of softplus(dt + bias) -> selective_scan (+ D*u) -> * silu(z) and its
    autograd gradients.  `t` holds the (possibly bf16-rounded) inputs; everything is widened to fp32 first."""
    leaf = {k: v.detach().float().cuda().requires_grad_(True) for k, v in t.items()}
    d = F.softplus(leaf["delta_raw"] + leaf["bias"])
    y = om.selective_scan(leaf["u"], d, leaf["A"], leaf["B"], leaf["C"], leaf["D"], impl="unbind")
    y = y * F.silu(leaf["z"])
    y.backward(dout.float().cuda())
    out = y.detach()
    grads = {k: v.grad for k, v in leaf.items()}
    del y, d
    return out, grads


def _product(t, dout, variant_bwd=0, chunk=None):
    from mamba_b200 import ops
    g = {k: v.detach().clone().cuda().requires_grad_(True) for k, v in t.items()}
    ops.SCAN_BWD_VARIANT = variant_bwd
    try:
        out = ops.selective_scan_fn(g["u"], g["delta_raw"], g["A"], g["B"], g["C"], g["D"], z=g["z"],
                                    delta_bias=g["bias"], delta_softplus=True, chunk=chunk)
        out.backward(dout.cuda())
    finally:
        ops.SCAN_BWD_VARIANT = 0
    return out.detach(), {k: v.grad for k, v in g.items()}


def _compare(out, grads, ref, rgrads, rtol, floor, what):
    assert_close(out, ref, rtol, floor, what=f"{what} out")
    for k in ("u", "delta_raw", "z", "B", "C"):
        assert_close(grads[k], rgrads[k], rtol, floor, what=f"{what} d{k}")
    # parameter gradients are sums over all B*L timesteps of signed terms: twice the floor
    for k in ("A", "D", "bias"):
        assert_close(grads[k], rgrads[k], rtol, 2 * floor, what=f"{what} d{k}")


def test_training_shape_bf16_fused_backward_vs_oracle():
    """(a) scan_bwd_fused_kernel<bf16, 8, 16> at the shape of the bench line: 129 chunks x 64 channel tiles."""
    B, L, D, N = 2, 2054, 2048, 64
    t = scan_inputs(B, L, D, N, seed=21, dtype=torch.bfloat16)
    dout = torch.randn(B, L, D, generator=torch.Generator().manual_seed(3)).bfloat16()
    ref, rg = _oracle_on_gpu(t, dout)
    out, g = _product(t, dout)
    assert out.dtype == torch.bfloat16
    _compare(out, g, ref, rg, RTOL16, FLOOR16, "training shape bf16")


def test_training_shape_fp32_lane_channel_backward_vs_oracle():
    """(b) the same shape in fp32 through scan_bwd_kernel (lane <-> channel)."""
    B, L, D, N = 2, 2054, 2048, 64
    t = scan_inputs(B, L, D, N, seed=22)
    dout = torch.randn(B, L, D, generator=torch.Generator().manual_seed(4))
    ref, rg = _oracle_on_gpu(t, dout)
    out, g = _product(t, dout)
    _compare(out, g, ref, rg, RTOL32, 1e-5, "training shape fp32")


@pytest.mark.parametrize("chunk", [8, 16])
def test_config5_long_context_fp32_vs_oracle(chunk):
    """(c) BASELINE config 5: L=8192, d_state 16, fp32, forward and backward."""
    B, L, D, N = 2, 8192, 2048, 16
    t = scan_inputs(B, L, D, N, seed=23)
    dout = torch.randn(B, L, D, generator=torch.Generator().manual_seed(5))
    ref, rg = _oracle_on_gpu(t, dout)
    out, g = _product(t, dout, chunk=chunk)
    _compare(out, g, ref, rg, RTOL32, 1e-5, f"config 5 fp32 chunk {chunk}")


def test_config5_ragged_edges_fp32_vs_oracle():
    """Config-5 kernels on a ragged problem: L not a multiple of the 64-step stage, D not a multiple of 32."""
    B, L, D, N = 3, 1000 + 37, 32 * 5 + 12, 16
    t = scan_inputs(B, L, D, N, seed=24)
    dout = torch.randn(B, L, D, generator=torch.Generator().manual_seed(6))
    ref, rg = _oracle_on_gpu(t, dout)
    for chunk in (8, 16):
        out, g = _product(t, dout, chunk=chunk)
        _compare(out, g, ref, rg, RTOL32, 1e-5, f"config 5 ragged chunk {chunk}")


def _full_args(cls):
    from mamba_b200.configs import common as cc
    return cls(d_model=1024, n_layer=10, vocab_size=cc.vocab_size, d_state=64, expand=2, d_conv=4,
               pad_vocab_size_multiple=1, metadata_vocab_size=cc.metadata_vocab_size)


def test_full_model_bf16_autocast_step_loss_and_gradients_vs_fp32_oracle():
    """(d) The benchmarked configuration: Layout P, d_model 1024, 10 layers, d_state 64, B=2 x T=2048, bf16 autocast
    (fp32 residual stream and scan state) against the fp32 oracle on the same weights and batch.

    Stated tolerance: loss within 1e-2 relative.  Gradients are held to a YARDSTICK, not to a constant: the oracle
    ITSELF run under torch's bf16 autocast on the same weights and batch.  At this depth and at random init that run
    is 15-25 % (relative L2) away from the fp32 oracle on most parameters (measured, printed below) - ten layers of
    bf16 GEMMs on a loss of ~5e2 - so a fixed few-percent bound would test nothing.  Per parameter, with
    e = ||g - g_ref|| / ||g_ref|| and c = cosine(g, g_ref) against the fp32 oracle:
        e <= max(0.05, 1.7 * e_torch)      and      1 - c <= max(0.03, 2 * (1 - c_torch)),
    e_torch / c_torch being the same quantities for the oracle under autocast.  The factor 1.7 covers what the
    product rounds in addition to torch autocast: the mixer's intermediate activations (conv output, dt, gate, scan
    output) are STORED in bf16 between kernels where the oracle under autocast keeps them in fp32 (measured: 1.6x on
    the metadata embedding, 1.02-1.05x elsewhere).  Parameters whose reference gradient is numerically zero (below
    1e-6 of the largest gradient) are held to that same absolute bound."""
    from mamba_b200 import synthetic, train
    from mamba_b200.models.mamba import Mamba, ModelArgs
    torch.manual_seed(0)
    ref = om.Mamba(_full_args(om.ModelArgs), scan_impl="unbind").cuda()
    model = Mamba(_full_args(ModelArgs)).cuda()
    model.load_state_dict(ref.state_dict(), strict=True)
    src, trg, meta = (x.cuda() for x in synthetic.batch(2, 2048, seed=31))

    lr = ref(src, meta, checkpoint_layers=True)
    loss_r = train_ref.loss_fn(src, trg, lr)
    loss_r.backward()
    del lr
    g_ref = {n: p.grad.float().clone() for n, p in ref.named_parameters()}
    ref.zero_grad(set_to_none=True)
    torch.cuda.empty_cache()
    # the oracle under torch's own bf16 autocast: the yardstick for what bf16 GEMMs cost on this model
    with torch.autocast("cuda", dtype=torch.bfloat16):
        la = ref(src, meta, checkpoint_layers=True)
    train_ref.loss_fn(src, trg, la.float()).backward()
    del la
    g_auto = {n: p.grad.float().clone() for n, p in ref.named_parameters()}
    ref.zero_grad(set_to_none=True)
    torch.cuda.empty_cache()

    with torch.autocast("cuda", dtype=torch.bfloat16):
        lg = model(src, meta)
    loss_g = train.loss_fn(src, trg, lg)
    loss_g.backward()
    torch.cuda.synchronize()

    rel_loss = abs(loss_g.item() - loss_r.item()) / abs(loss_r.item())
    print(f"[fullsize] loss oracle {loss_r.item():.6f} product {loss_g.item():.6f} rel {rel_loss:.2e}")
    assert rel_loss <= 1e-2, (loss_g.item(), loss_r.item())
    gp = dict(model.named_parameters())
    gmax = max(float(g.abs().max()) for g in g_ref.values())
    worst = (0.0, 1.0, "", 0.0)
    for name, gr in g_ref.items():
        gg = gp[name].grad.float()
        if float(gr.abs().max()) < 1e-6 * gmax:
            assert float(gg.abs().max()) < 1e-4 * gmax, name
            continue
        rel = float((gg - gr).norm() / gr.norm())
        rel_torch = float((g_auto[name] - gr).norm() / gr.norm())
        cos = float(F.cosine_similarity(gg.flatten(), gr.flatten(), dim=0))
        cos_torch = float(F.cosine_similarity(g_auto[name].flatten(), gr.flatten(), dim=0))
        if rel > worst[0]:
            worst = (rel, cos, name, rel_torch)
        if rel > 0.03:
            print(f"[fullsize] {name}: rel-L2 {rel:.3e} (oracle under torch autocast {rel_torch:.3e}) cos {cos:.5f} ({cos_torch:.5f})")
        assert rel <= max(0.05, 1.7 * rel_torch) and 1 - cos <= max(0.03, 2 * (1 - cos_torch)), (name, rel, rel_torch, cos, cos_torch)
    print(f"[fullsize] worst parameter gradient: {worst[2]} rel-L2 {worst[0]:.3e} (oracle under torch autocast: "
          f"{worst[3]:.3e}) cos {worst[1]:.5f}")


def test_greedy_decode_2000_tokens_recurrent_vs_literal_default_model():
    """(e) BASELINE config 4: 5 composer bands x 2 samples = 10 sequences, 2000 new tokens, default model
    (scripts/generate_midi_combined.py:60-127 drives scripts/generate_midi_many.py:13-56).  The recurrent decoder
    (one CUDA-graphed step per token, running logsumexp over the sequence axis) must emit the same tokens as the
    literal loop that re-runs the full model on the growing window.

    A row may part from the literal sequence only at a step where the literal loop's own top-2 margin is below
    1e-4 of its largest |filtered logit| (a numerical tie: the two paths sum in different orders); from there on
    that row's continuation is legitimately different and is no longer compared.  The count of such rows and
    the number of bit-identical tokens are printed; at least half of all tokens must have been compared equal."""
    from mamba_b200 import generate, synthetic, train
    from mamba_b200.configs import common as cc
    from mamba_b200.models.mamba import Mamba, ModelArgs
    torch.manual_seed(0)
    model = Mamba(_full_args(ModelArgs)).cuda().eval()
    nseq, prompt, n_new = 10, 48, 2000
    src, _, meta = synthetic.batch(nseq, prompt, seed=41)
    src, meta = src.cuda(), meta.cuda()
    rec = generate.generate_recurrent(model, src, meta, n_new, use_graph=True)[:, prompt:]

    # literal loop (generate.generate_literal) with the top-2 margin of every choice recorded
    table = generate._penalty_table_many(src.device)
    token_ids, generated = src.clone(), src.clone()
    lit = torch.empty(nseq, n_new, dtype=torch.long, device=src.device)
    margin = torch.empty(nseq, n_new, device=src.device)
    with torch.no_grad():
        for i in range(n_new):
            logits = model(token_ids, meta)
            last = train.filtered_logit(token_ids, logits)[:, -1, :]
            recent = generated[:, -100:]
            counts = torch.zeros(nseq, cc.vocab_size, device=src.device).scatter_add_(
                1, recent, torch.ones_like(recent, dtype=torch.float32))
            last = generate._apply_penalty_many(last, counts, table)
            top2 = last.topk(2, dim=-1).values
            margin[:, i] = (top2[:, 0] - top2[:, 1]) / last.abs().max(dim=-1).values
            nxt = last.argmax(-1, keepdim=True)
            lit[:, i] = nxt[:, 0]
            generated = torch.cat([generated, nxt], dim=1)
            token_ids = torch.cat([token_ids, nxt], dim=1)
    lit, rec, margin = lit.cpu(), rec.cpu(), margin.cpu()
    equal_tokens, tied_rows = 0, 0
    for b in range(nseq):
        diff = (lit[b] != rec[b]).nonzero()
        if diff.numel() == 0:
            equal_tokens += n_new
            continue
        first = int(diff[0])
        equal_tokens += first
        tied_rows += 1
        assert float(margin[b, first]) < 1e-4, (
            f"row {b} parts from the literal loop at step {first} with top-2 margin {float(margin[b, first]):.3e}")
    print(f"[fullsize] greedy decode: {equal_tokens}/{nseq * n_new} tokens bit-identical, {tied_rows} rows left at a tie")
    assert equal_tokens >= nseq * n_new // 2
