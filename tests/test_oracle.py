"""CPU tests of the oracle itself (no GPU): internal consistency of the restatement."""
import torch
import torch.nn.functional as F

from oracle import simple_mamba as om
from oracle import train_ref


def _args(**kw):
    base = dict(d_model=32, n_layer=2, vocab_size=17914, d_state=8, expand=2, d_conv=4, metadata_vocab_size=568,
                pad_vocab_size_multiple=1)
    base.update(kw)
    return om.ModelArgs(**base)


def test_literal_equals_unbind_bitwise_forward_and_grads():
    """the `unbind` form is bit-identical to the literal loop (@L325-328), fwd and all grads."""
    torch.manual_seed(0)
    blocks = {}
    grads = {}
    outs = {}
    x0 = torch.randn(2, 24, 32)
    for impl in ("literal", "unbind"):
        torch.manual_seed(1)
        blk = om.MambaBlock(_args(), scan_impl=impl)
        x = x0.clone().requires_grad_(True)
        y = blk(x)
        y.square().sum().backward()
        outs[impl] = y.detach()
        grads[impl] = [x.grad] + [p.grad for p in blk.parameters()]
        blocks[impl] = blk
    assert torch.equal(outs["literal"], outs["unbind"])
    for a, b in zip(grads["literal"], grads["unbind"]):
        assert torch.equal(a, b)


def test_step_matches_full_forward_last_position():
    """the single-token recurrence reproduces the last position of the full forward."""
    torch.manual_seed(0)
    args = _args()
    blk = om.MambaBlock(args).eval()
    x = torch.randn(2, 17, 32)
    with torch.no_grad():
        full = blk(x)
        conv_state = torch.zeros(2, args.d_inner, args.d_conv)
        ssm_state = torch.zeros(2, args.d_inner, args.d_state)
        for t in range(x.shape[1]):
            y, conv_state, ssm_state = blk.step(x[:, t], conv_state, ssm_state)
            assert (y - full[:, t]).abs().max() < 1e-5


def test_param_count_layout_p_at_config():
    """88,554,496 parameters at the configured sizes with the padded vocabulary."""
    a = om.ModelArgs(d_model=1024, n_layer=10, vocab_size=17914, d_state=64, expand=2, d_conv=4)
    assert a.vocab_size == 17920 and a.d_inner == 2048 and a.dt_rank == 64
    per_block = (4096 * 1024 + 2048 * 4 + 2048 + 192 * 2048 + 2048 * 64 + 2048 + 2048 * 64 + 2048 + 1024 * 2048)
    assert per_block == 6961152
    total = 10 * (per_block + 1024) + 17920 * 1024 + 568 * 1024 + 1024
    assert total == 88554496


def test_softplus_threshold_and_scan_shapes():
    u = torch.randn(1, 5, 4)
    delta = F.softplus(torch.tensor([[[-3.0, 0.0, 19.9, 25.0]]]).expand(1, 5, 4))
    assert delta[0, 0, 3] == 25.0  # identity above the threshold of 20
    A = -torch.ones(4, 3)
    y, h = om.selective_scan(u, delta, A, torch.randn(1, 5, 3), torch.randn(1, 5, 3), torch.ones(4),
                             return_last_state=True)
    assert y.shape == (1, 5, 4) and h.shape == (1, 4, 3)


def test_distributions_table():
    """train.py:79-111: rows are indexed by the bucket of the PREVIOUS token."""
    V, s = train_ref.vocab_layout()
    assert V == 17914 and s == {"pitch": 0, "dyn": 16512, "length": 16640, "time": 17152, "tempo": 17664}
    d = train_ref.make_distributions()
    assert d.shape == (5, V)
    assert d[0, s["dyn"]] == 1 and d[0, s["length"] - 1] == 0          # after a pitch token: dyn (last id excluded)
    assert d[4, 0] == 10 and d[4, s["dyn"] - 1] == 0                   # after a tempo token: pitch x10
    assert d[1, s["length"]] == 1 and abs(float(d[1, s["time"] - 2]) - 3.0) < 1e-6   # length ramp 1..3
    assert d[2, s["time"]] == 1 and d[2, s["tempo"]] == 1              # after length: time or tempo
    assert d[3, s["tempo"]] == 1 and d[3, V - 1] == 1


def test_filtered_logit_softmax_axis_is_sequence():
    """log_softmax(dim=1) normalises over positions, not vocab."""
    V, _ = train_ref.vocab_layout()
    torch.manual_seed(0)
    src = torch.randint(0, V, (2, 7))
    out = torch.randn(2, 7, V)
    f = train_ref.filtered_logit(src, out)
    w = train_ref.pick_distributions_by_prev_token(src)
    lp = out - torch.logsumexp(out, dim=1, keepdim=True)
    assert torch.allclose(f, -lp * w, atol=1e-6)
