"""CPU tests of the host-side mirror of the reference interface (configs, loss path, synthetic data,
module construction / state_dict layout)."""
import torch

from mamba_b200 import synthetic, train
from mamba_b200.configs import common as cc
from mamba_b200.models.mamba import Mamba, MambaBlock, ModelArgs
from oracle import simple_mamba as om
from oracle import train_ref


def test_config_values_match_reference():
    V, s = train_ref.vocab_layout()
    assert cc.vocab_size == V == 17914
    assert cc.start_idx == s
    assert cc.metadata_vocab_size == 568
    assert cc.config.values.block_len == 2048 and cc.config.values.batch_size == 2
    assert cc.config.values.learning_rate == 5e-5
    d = train.get_mamba_dict()
    assert (d.d_model, d.n_layer, d.d_state, d.expand, d.d_conv, d.d_inner, d.dt_rank) == (1024, 10, 64, 2, 4, 2048, 64)
    assert train.get_mamba_dict(pad_vocab=True).vocab_size == 17920 == train.get_actual_vocab_size("mamba")


def test_loss_path_matches_oracle_on_cpu():
    torch.manual_seed(0)
    src, trg, _ = synthetic.batch(2, 32, seed=3)
    out = torch.randn(2, 32, cc.vocab_size)
    assert torch.equal(train.make_distributions("cpu"), train_ref.make_distributions())
    assert torch.equal(train.pick_distributions_by_prev_token(src), train_ref.pick_distributions_by_prev_token(src))
    assert torch.equal(train.filtered_logit(src, out), train_ref.filtered_logit(src, out))
    assert torch.equal(train.loss_fn_torch(src, trg, out), train_ref.loss_fn(src, trg, out))


def test_synthetic_batch_follows_grammar():
    src, trg, meta = synthetic.batch(3, 50, seed=1)
    assert src.shape == trg.shape == (3, 50) and meta.shape == (3, 6)
    assert torch.equal(src[:, 1:], trg[:, :-1])
    s = cc.start_idx
    bounds = [s["pitch"], s["dyn"], s["length"], s["time"], s["tempo"], cc.vocab_size]
    for p in range(50):
        c = p % 5
        assert (src[:, p] >= bounds[c]).all() and (src[:, p] < bounds[c + 1]).all()
    assert (meta[:, 0] >= 313).all() and (meta[:, 0] <= 567).all()
    assert (meta[:, 1:5] >= 203).all() and (meta[:, 1:5] <= 311).all()
    assert (meta[:, 5] >= 1).all() and (meta[:, 5] <= 201).all()
    again = synthetic.batch(3, 50, seed=1)
    assert torch.equal(src, again[0]) and torch.equal(meta, again[2])


def _small_args(cls):
    return cls(d_model=32, n_layer=2, vocab_size=cc.vocab_size, d_state=8, expand=2, d_conv=4,
               pad_vocab_size_multiple=1, metadata_vocab_size=cc.metadata_vocab_size)


def test_state_dict_layout_matches_oracle_layout_p():
    prod = Mamba(_small_args(ModelArgs))
    ref = om.Mamba(_small_args(om.ModelArgs))
    sp, sr = prod.state_dict(), ref.state_dict()
    assert list(sp.keys()) == list(sr.keys())
    for k in sp:
        assert sp[k].shape == sr[k].shape and sp[k].dtype == sr[k].dtype, k
    ref.load_state_dict(sp, strict=True)
    prod.load_state_dict(ref.state_dict(), strict=True)
    assert prod.lm_head.weight is prod.embedding.weight  # tied (simple_mamba @L70)


def test_state_dict_layout_shipped_wrapper():
    """Layout S = models/mamba/mamba.py:8-35 with mamba_ssm.Mamba2 layers: outer keys, Mamba-2 parameter names and
    shapes, identical to the oracle restatement, and loadable both ways."""
    from oracle.mamba2_ref import ShippedMambaRef
    m = Mamba(d_model=64, n_layers=2)
    keys = list(m.state_dict().keys())
    assert keys[:4] == ["token_embedding.weight", "metadata_embedding.weight", "output_layer.weight",
                        "output_layer.bias"]
    assert keys[-2:] == ["norm.weight", "norm.bias"]
    layer = {k.split(".", 2)[2]: tuple(v.shape) for k, v in m.state_dict().items() if k.startswith("layers.1.")}
    d_inner, N, H = 128, 64, 2
    assert layer == {"dt_bias": (H,), "A_log": (H,), "D": (H,), "in_proj.weight": (2 * d_inner + 2 * N + H, 64),
                     "conv1d.weight": (d_inner + 2 * N, 1, 4), "conv1d.bias": (d_inner + 2 * N,),
                     "norm.weight": (d_inner,), "out_proj.weight": (64, d_inner)}
    ref = ShippedMambaRef(d_model=64, n_layers=2, d_state=64)
    assert sorted(ref.state_dict().keys()) == sorted(keys)
    m.load_state_dict(ref.state_dict(), strict=True)
    ref.load_state_dict(m.state_dict(), strict=True)
    assert m.token_embedding.weight.shape == (17914, 64)


def test_shipped_model_parameter_count_is_the_number_the_reference_prints():
    """scripts/Test Accuracy.ipynb:52 prints 101,972,666 trainable parameters for Mamba(d_model=1024, n_layers=10):
    the one quantitative anchor the reference gives for the shipped (Mamba-2) layout."""
    m = Mamba(d_model=1024, n_layers=10)
    assert sum(p.numel() for p in m.parameters() if p.requires_grad) == 101_972_666


def test_block_init_matches_reference_init():
    blk = MambaBlock(_small_args(ModelArgs))
    assert torch.equal(blk.A_log, torch.log(torch.arange(1, 9, dtype=torch.float32)).repeat(64, 1))
    assert torch.equal(blk.D, torch.ones(64))
    assert blk.conv1d.weight.shape == (64, 1, 4) and blk.x_proj.weight.shape == (2 + 16, 64)
    assert blk.in_proj.bias is None and blk.dt_proj.bias is not None


def test_default_model_param_count():
    m = train.new_model("mamba")
    assert sum(p.numel() for p in m.parameters()) == 88554496 - 6 * 1024  # unpadded vocabulary


def test_split_fn_backward_uses_the_arena_or_falls_back_to_cat():
    """ops.SplitFn (pure torch): when every part's gradient already lives in its slice of the plan's arena the
    backward returns the arena itself (no concatenation); any other gradient takes the ordinary cat path."""
    from mamba_b200 import ops
    torch.manual_seed(0)
    x = torch.randn(2, 5, 7, requires_grad=True)
    # (1) gradients produced elsewhere -> cat fallback, equal to plain split()
    plan = ops.MixerGradPlan(x.shape, None)
    a, b = ops.split_fn(x, [3, 4], plan, "xz")
    ga, gb = torch.randn(2, 5, 3), torch.randn(2, 5, 4)
    (gx,) = torch.autograd.grad([a, b], [x], [ga, gb])
    assert torch.equal(gx, torch.cat([ga, gb], dim=-1))
    # (2) gradients that ARE the arena's slices -> the arena comes back untouched
    plan = ops.MixerGradPlan(x.shape, None)
    a, b = ops.split_fn(x, [3, 4], plan, "xz")
    pa = plan.part("xz", x.shape, x.dtype, x.device, 0, 3)
    pb = plan.part("xz", x.shape, x.dtype, x.device, 3, 4)
    pa.copy_(ga), pb.copy_(gb)
    (gx,) = torch.autograd.grad([a, b], [x], [pa, pb])
    assert gx.data_ptr() == plan.arena["xz"].data_ptr() and torch.equal(gx, torch.cat([ga, gb], dim=-1))
    # (3) one part missing (None gradient) -> zeros for it
    plan = ops.MixerGradPlan(x.shape, None)
    a, b = ops.split_fn(x, [3, 4], plan, "xz")
    (gx,) = torch.autograd.grad([a], [x], [ga])
    assert torch.equal(gx, torch.cat([ga, torch.zeros(2, 5, 4)], dim=-1))
    # without a plan it is torch's own split
    a, b = ops.split_fn(x, [3, 4])
    assert a.shape == (2, 5, 3) and b.shape == (2, 5, 4)


def test_weight_shadows_are_used_only_while_the_parameter_is_unchanged():
    """models.mamba.WeightShadows: a cast copy is handed out only while the parameter's version counter is the one it
    was made from; any other in-place writer (load_state_dict, optimizer of another owner) makes it fall back."""
    from mamba_b200.models.mamba.mamba import WeightShadows
    w = torch.nn.Parameter(torch.randn(4, 6))
    other = torch.nn.Parameter(torch.randn(4, 6))
    keep = WeightShadows.register([w], torch.bfloat16)
    sh = WeightShadows.get(w, torch.bfloat16)
    assert sh is keep[0] and torch.equal(sh, w.detach().to(torch.bfloat16))
    assert WeightShadows.get(w, torch.float16) is None and WeightShadows.get(other, torch.bfloat16) is None
    with torch.no_grad():
        w.mul_(2.0)                                   # someone else writes the parameter
    assert WeightShadows.stale([w]) and WeightShadows.get(w, torch.bfloat16) is None
    WeightShadows.refresh([w])
    assert not WeightShadows.stale([w])
    assert torch.equal(WeightShadows.get(w, torch.bfloat16), w.detach().to(torch.bfloat16))
    WeightShadows.table.pop(id(w))


def test_mamba2_oracle_recurrence_equals_the_published_dual_form():
    """oracle/mamba2_ref.py cannot be pinned against mamba_ssm (absent).  What can be checked without it: the two
    PUBLISHED forms of the Mamba-2 operator agree — the recurrence the oracle restates, and the state-space-dual
    (masked-attention) form of the same paper, written here independently:
        Y = (Lmask * (C B^T)) (dt * X) + D X,   Lmask[i, j] = exp(sum_{k=j+1..i} dt_k A) for i >= j, else 0."""
    import torch.nn.functional as F
    from oracle.mamba2_ref import Mamba2Ref
    torch.manual_seed(3)
    m = Mamba2Ref(d_model=32, d_state=16, d_conv=4, expand=2, headdim=8).double()
    with torch.no_grad():
        m.D.add_(0.3 * torch.randn_like(m.D))
        m.norm.weight.add_(0.1 * torch.randn_like(m.norm.weight))
    u = torch.randn(2, 19, 32, dtype=torch.double)
    with torch.no_grad():
        want = m(u)
        # dual form, head by head
        Bsz, L, _ = u.shape
        H, P, N = m.nheads, m.headdim, m.d_state
        z, xBC, dt = torch.split(m.in_proj(u), [m.d_inner, m.d_inner + 2 * N, H], dim=-1)
        dt = F.softplus(dt + m.dt_bias)
        xBC = F.silu(m.conv1d(xBC.transpose(1, 2))[..., :L].transpose(1, 2))
        x, Bm, Cm = torch.split(xBC, [m.d_inner, N, N], dim=-1)
        x = x.reshape(Bsz, L, H, P)
        A = -torch.exp(m.A_log)
        cum = torch.cumsum(dt * A, dim=1)                                        # [B, L, H]
        seg = cum[:, :, None, :] - cum[:, None, :, :]                            # [B, i, j, H] = sum_{k=j+1..i}
        mask = torch.tril(torch.ones(L, L, dtype=torch.bool))[None, :, :, None]
        Lm = torch.where(mask, torch.exp(seg.masked_fill(~mask, 0.0)), torch.zeros((), dtype=torch.double))
        G = torch.einsum("bin,bjn->bij", Cm, Bm)                                 # C B^T (one group)
        y = torch.einsum("bij,bijh,bjh,bjhp->bihp", G, Lm, dt, x) + m.D[None, None, :, None] * x
        y = y.reshape(Bsz, L, m.d_inner) * F.silu(z)
        y = y * torch.rsqrt(y.pow(2).mean(-1, keepdim=True) + m.norm_eps) * m.norm.weight
        got = m.out_proj(y)
    # (the oracle carries its state in fp32 whatever the module's dtype: agreement to fp32 rounding)
    assert torch.allclose(got, want, rtol=1e-5, atol=2e-6), float((got - want).abs().max())
