"""torch.autograd.Function wrappers over the C-ABI (include/mamba_b200.h).

PyTorch is plumbing here: it owns device memory, the current stream and autograd bookkeeping; all
arithmetic of the hot path runs in libmamba_b200.so.  Nothing in this module computes on the CPU
and nothing falls back to torch ops: a missing library or a non-CUDA tensor raises.

Reference behaviour replaced (models/mamba/__pycache__/simple_mamba.cpython-311.pyc, see SURVEY.md
Appendix A):  selective_scan_fn -> MambaBlock.selective_scan @L310-333 (+ softplus @L276, D skip
@L331, gate @L241);  causal_conv1d_silu_fn -> @L233-237;  rmsnorm_fn -> RMSNorm.forward @L346 and
the residual add of ResidualBlock.forward @L179;  conv_step / ssm_step -> the same maths for one
new token (no reference counterpart: scripts/generate.py:26-31 re-runs the full model).
"""
from __future__ import annotations

import ctypes as ct
import os

import torch

from . import _lib
from ._lib import (FLAG_A_IS_LOG, FLAG_DELTA_SOFTPLUS, FLAG_HAS_D, FLAG_HAS_DELTA_BIAS, FLAG_HAS_Z, MAMBA_BF16, MAMBA_F32, ConvArgs,
                   DecodeLayer, DecodeTokenArgs, FusedLinearStepArgs, LinearStepArgs, SampleStepArgs, LossArgs, NormArgs, ScanBwdArgs, ScanFwdArgs, StepArgs, check, lib)

_DTYPES = {torch.float32: MAMBA_F32, torch.bfloat16: MAMBA_BF16}

# tuning knobs (0 = library default); exposed for the kernel sweeps in bench.py
SCAN_CHUNK = int(os.environ.get("MAMBA_B200_SCAN_CHUNK", "16"))
SCAN_FWD_VARIANT = int(os.environ.get("MAMBA_B200_FWD_VARIANT", "0"))
SCAN_BWD_VARIANT = int(os.environ.get("MAMBA_B200_BWD_VARIANT", "0"))


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "mamba_b200 ops run only on CUDA tensors (sm_100a kernels); there is no CPU fallback — "
                f"got a tensor on {t.device}")


def _dtype_code(t: torch.Tensor) -> int:
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise TypeError(f"mamba_b200: unsupported activation dtype {t.dtype} (float32 or bfloat16)") from None


def _rows(t: torch.Tensor) -> torch.Tensor:
    """[B, L, X] tensor with unit stride on the last axis (views of split() stay views)."""
    return t if t.stride(-1) == 1 else t.contiguous()


def _p(t):
    return None if t is None else ct.c_void_p(t.data_ptr())


def _stream() -> ct.c_void_p:
    return ct.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32(t):
    if t is None:
        return None
    return t.detach().to(torch.float32).contiguous()


# Optional per-op device timing (bench.py's roofline leg): when KERNEL_TIMES is a dict, every C-ABI call is
# bracketed by CUDA events on the launching stream and the (start, end) pairs are collected per entry point.
KERNEL_TIMES = None


def _call(name, a, dev):
    with torch.cuda.device(dev):
        if KERNEL_TIMES is None:
            check(getattr(lib(), name)(ct.byref(a), _stream()), name)
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(getattr(lib(), name)(ct.byref(a), _stream()), name)
        e1.record()
        KERNEL_TIMES.setdefault(name, []).append((e0, e1))


# ------------------------------------------------------------------------------------------------
# gradient arenas for the mixer's split() views
# ------------------------------------------------------------------------------------------------
class MixerGradPlan:
    """One per MambaBlock.forward call.  The mixer splits two GEMM outputs into views (xz -> x | z, x_dbl ->
    dt | B | C); autograd's backward of a split is a concatenation of the parts' gradients.  With a plan the kernels
    that PRODUCE those gradients write them straight into the right slice of one buffer per split (the C-ABI takes
    strided outputs), and split's backward hands that buffer on without a copy.  Everything falls back to the
    ordinary concatenation if a gradient did not come from the arena."""

    def __init__(self, xz_shape=None, xdbl_shape=None):
        self.arena = {}
        self.xz_shape = None if xz_shape is None else tuple(xz_shape)        # [B, L, 2 * d_inner]
        self.xdbl_shape = None if xdbl_shape is None else tuple(xdbl_shape)  # [B, L, dt_rank + 2 * d_state]

    def buffer(self, key, shape, dtype, device):
        b = self.arena.get(key)
        if b is None or b.shape != tuple(shape) or b.dtype != dtype:
            b = torch.empty(tuple(shape), dtype=dtype, device=device)
            self.arena[key] = b
        return b

    def part(self, key, shape, dtype, device, start, width):
        """The [..., start:start+width] slice of arena `key` (allocated on first use)."""
        return self.buffer(key, shape, dtype, device)[..., start:start + width]


class SplitFn(torch.autograd.Function):
    """x.split(sizes, dim=-1) whose backward returns the plan's arena when every part's gradient already lives in
    its slice of it (no concatenation kernel); otherwise the usual cat."""

    @staticmethod
    def forward(ctx, x, plan, key, *sizes):
        ctx.plan, ctx.key, ctx.sizes = plan, key, sizes
        ctx.xshape, ctx.xdtype = x.shape, x.dtype
        return tuple(v.view_as(v) for v in x.split(list(sizes), dim=-1))

    @staticmethod
    def backward(ctx, *grads):
        buf = ctx.plan.arena.get(ctx.key) if ctx.plan is not None else None
        ok = buf is not None and buf.shape == ctx.xshape and buf.dtype == ctx.xdtype
        off = 0
        for g, w in zip(grads, ctx.sizes):
            if ok:
                want = buf[..., off:off + w]
                ok = (g is not None and g.dtype == buf.dtype and g.shape == want.shape and g.stride() == want.stride()
                      and g.data_ptr() == want.data_ptr())
            off += w
        if ok:
            return (buf, None, None) + (None,) * len(ctx.sizes)
        parts = [g if g is not None else torch.zeros(ctx.xshape[:-1] + (w,), dtype=ctx.xdtype, device=grads[0].device if grads[0] is not None else None)
                 for g, w in zip(grads, ctx.sizes)]
        return (torch.cat([p_.to(ctx.xdtype) for p_ in parts], dim=-1), None, None) + (None,) * len(ctx.sizes)


def split_fn(x, sizes, plan=None, key=None):
    if plan is None:
        return x.split(list(sizes), dim=-1)
    return SplitFn.apply(x, plan, key, *sizes)


# ------------------------------------------------------------------------------------------------
# selective scan
# ------------------------------------------------------------------------------------------------
def _scan_fwd_raw(u, delta, A, B, C, D, z, delta_bias, delta_softplus, ckpt, chunk, h_init=None, h_last=None,
                  variant=0, y_pre=None, out=None, a_is_log=False):
    Bsz, L, Dm = u.shape
    N = A.shape[1]
    if out is None:
        out = torch.empty((Bsz, L, Dm), dtype=u.dtype, device=u.device)
    a = ScanFwdArgs()
    a.struct_size = ct.sizeof(ScanFwdArgs)
    a.dtype = _dtype_code(u)
    a.batch, a.seqlen, a.dim, a.dstate = Bsz, L, Dm, N
    a.chunk = chunk
    a.flags = ((FLAG_HAS_Z if z is not None else 0) | (FLAG_DELTA_SOFTPLUS if delta_softplus else 0)
               | (FLAG_HAS_DELTA_BIAS if delta_bias is not None else 0) | (FLAG_HAS_D if D is not None else 0)
               | (FLAG_A_IS_LOG if a_is_log else 0))
    a.variant = variant
    a.u, a.u_bs, a.u_ls = _p(u), u.stride(0), u.stride(1)
    a.delta, a.delta_bs, a.delta_ls = _p(delta), delta.stride(0), delta.stride(1)
    a.A = _p(A)
    a.B, a.B_bs, a.B_ls = _p(B), B.stride(0), B.stride(1)
    a.C, a.C_bs, a.C_ls = _p(C), C.stride(0), C.stride(1)
    a.D = _p(D)
    if z is not None:
        a.z, a.z_bs, a.z_ls = _p(z), z.stride(0), z.stride(1)
    a.delta_bias = _p(delta_bias)
    a.out, a.out_bs, a.out_ls = _p(out), out.stride(0), out.stride(1)
    a.ckpt = _p(ckpt)
    a.h_last = _p(h_last)
    a.h_init = _p(h_init)
    if y_pre is not None:
        a.y_pre, a.y_pre_bs, a.y_pre_ls = _p(y_pre), y_pre.stride(0), y_pre.stride(1)
    _call("mamba_scan_fwd", a, u.device)
    return out


class SelectiveScanFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u, delta, A, B, C, D, z, delta_bias, delta_softplus, chunk, a_is_log=False, plan=None):
        _require_cuda(u, delta, A, B, C, D, z, delta_bias)
        ctx.plan = plan
        if u.dim() != 3 or delta.shape != u.shape:
            raise ValueError(f"selective_scan: u {tuple(u.shape)} and delta {tuple(delta.shape)} must be equal [B, L, D]")
        if A.dim() != 2 or A.shape[0] != u.shape[2]:
            raise ValueError(f"selective_scan: A {tuple(A.shape)} must be [D={u.shape[2]}, N]")
        if B.shape != (u.shape[0], u.shape[1], A.shape[1]) or C.shape != B.shape:
            raise ValueError(f"selective_scan: B {tuple(B.shape)} / C {tuple(C.shape)} must be [B, L, N]")
        dt = u.dtype
        _dtype_code(u)
        u, delta, B, C = (_rows(t.to(dt)) for t in (u, delta, B, C))
        z = None if z is None else _rows(z.to(dt))
        A32, D32, b32 = _f32(A), _f32(D), _f32(delta_bias)
        need_grad = any(ctx.needs_input_grad[:8])
        ckpt = y_pre = None
        if need_grad and u.shape[1] > chunk:
            n = lib().mamba_scan_ckpt_elems(u.shape[0], u.shape[1], u.shape[2], A.shape[1], chunk)
            ckpt = torch.empty(n, dtype=torch.float32, device=u.device)
        if need_grad and z is not None:
            y_pre = torch.empty(u.shape, dtype=u.dtype, device=u.device)  # pre-gate output, for dz
        out = _scan_fwd_raw(u, delta, A32, B, C, D32, z, b32, delta_softplus, ckpt, chunk, variant=SCAN_FWD_VARIANT,
                            y_pre=y_pre, a_is_log=a_is_log)
        ctx.save_for_backward(u, delta, A32, B, C, D32, z, b32, ckpt, y_pre)
        ctx.delta_softplus = delta_softplus
        ctx.a_is_log = a_is_log
        ctx.chunk = chunk
        ctx.in_dtypes = (A.dtype, None if D is None else D.dtype, None if delta_bias is None else delta_bias.dtype)
        return out

    @staticmethod
    def backward(ctx, dout):
        u, delta, A, B, C, D, z, dbias, ckpt, y_pre = ctx.saved_tensors
        Bsz, L, Dm = u.shape
        N = A.shape[1]
        dout = _rows(dout.to(u.dtype))
        dev = u.device
        du = torch.empty_like(u, memory_format=torch.contiguous_format)
        ddelta = torch.empty((Bsz, L, Dm), dtype=u.dtype, device=dev)
        plan = ctx.plan
        if plan is not None and z is not None and plan.xz_shape is not None and plan.xdbl_shape is not None:
            # gradients of the split() views go straight into their slices of the arenas (see MixerGradPlan)
            dz = plan.part("xz", plan.xz_shape, u.dtype, dev, Dm, Dm)
            R = plan.xdbl_shape[-1] - 2 * N
            dB = plan.part("xdbl", plan.xdbl_shape, u.dtype, dev, R, N)
            dC = plan.part("xdbl", plan.xdbl_shape, u.dtype, dev, R + N, N)
        else:
            dz = torch.empty((Bsz, L, Dm), dtype=u.dtype, device=dev) if z is not None else None
            dB = torch.empty((Bsz, L, N), dtype=u.dtype, device=dev)
            dC = torch.empty((Bsz, L, N), dtype=u.dtype, device=dev)
        dA = torch.empty((Dm, N), dtype=torch.float32, device=dev)
        dD = torch.empty((Dm,), dtype=torch.float32, device=dev) if D is not None else None
        ddb = torch.empty((Dm,), dtype=torch.float32, device=dev) if dbias is not None else None
        wsb = lib().mamba_scan_bwd_workspace_bytes(Bsz, L, Dm, N)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        a = ScanBwdArgs()
        a.struct_size = ct.sizeof(ScanBwdArgs)
        a.dtype = _dtype_code(u)
        a.batch, a.seqlen, a.dim, a.dstate = Bsz, L, Dm, N
        a.chunk = ctx.chunk
        a.flags = ((FLAG_HAS_Z if z is not None else 0) | (FLAG_DELTA_SOFTPLUS if ctx.delta_softplus else 0)
                   | (FLAG_HAS_DELTA_BIAS if dbias is not None else 0) | (FLAG_HAS_D if D is not None else 0)
                   | (FLAG_A_IS_LOG if ctx.a_is_log else 0))
        a.variant = SCAN_BWD_VARIANT
        a.u, a.u_bs, a.u_ls = _p(u), u.stride(0), u.stride(1)
        a.delta, a.delta_bs, a.delta_ls = _p(delta), delta.stride(0), delta.stride(1)
        a.A = _p(A)
        a.B, a.B_bs, a.B_ls = _p(B), B.stride(0), B.stride(1)
        a.C, a.C_bs, a.C_ls = _p(C), C.stride(0), C.stride(1)
        a.D = _p(D)
        if z is not None:
            a.z, a.z_bs, a.z_ls = _p(z), z.stride(0), z.stride(1)
            a.dz, a.dz_bs, a.dz_ls = _p(dz), dz.stride(0), dz.stride(1)
            a.y_pre, a.y_pre_bs, a.y_pre_ls = _p(y_pre), y_pre.stride(0), y_pre.stride(1)
        a.delta_bias = _p(dbias)
        a.dout, a.dout_bs, a.dout_ls = _p(dout), dout.stride(0), dout.stride(1)
        a.ckpt = _p(ckpt)
        a.du, a.du_bs, a.du_ls = _p(du), du.stride(0), du.stride(1)
        a.ddelta, a.ddelta_bs, a.ddelta_ls = _p(ddelta), ddelta.stride(0), ddelta.stride(1)
        a.dB, a.dB_bs, a.dB_ls = _p(dB), dB.stride(0), dB.stride(1)
        a.dC, a.dC_bs, a.dC_ls = _p(dC), dC.stride(0), dC.stride(1)
        a.dA, a.dD, a.ddelta_bias = _p(dA), _p(dD), _p(ddb)
        a.workspace, a.workspace_bytes = _p(ws), wsb
        _call("mamba_scan_bwd", a, dev)
        tA, tD, tb = ctx.in_dtypes
        return (du, ddelta, dA.to(tA), dB, dC, None if dD is None else dD.to(tD), dz,
                None if ddb is None else ddb.to(tb), None, None, None, None)




def selective_scan_fn(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False, chunk=None,
                      A_is_log=False, plan=None):
    """Fused selective scan.  u, delta: [B, L, D]; A: [D, N]; B, C: [B, L, N]; D, delta_bias: [D];
    z: [B, L, D].  Returns out [B, L, D] = (scan(u, delta, A, B, C) + D*u) * silu(z).
    A_is_log: `A` is the A_log parameter; the kernels form A = -exp(A_log) (simple_mamba @L270) themselves and the
    gradient comes back w.r.t. A_log (saves five elementwise launches per layer and step)."""
    chunk = SCAN_CHUNK if chunk is None else int(chunk)
    if A.shape[1] > 64:
        chunk = min(chunk, 8)   # the backward's shared-memory tiles for d_state > 64 only fit 8-step chunks (csrc/scan_bwd.cu)
    return SelectiveScanFn.apply(u, delta, A, B, C, D, z, delta_bias, bool(delta_softplus), chunk, bool(A_is_log), plan)


def selective_scan_prefill(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False, h_init=None):
    """Inference-only scan that also returns the final state h_{L-1} ([B, D, N] fp32) for decode."""
    _require_cuda(u, delta, A, B, C)
    dt = u.dtype
    u, delta, B, C = (_rows(t.to(dt)) for t in (u, delta, B, C))
    z = None if z is None else _rows(z.to(dt))
    h_last = torch.empty((u.shape[0], u.shape[2], A.shape[1]), dtype=torch.float32, device=u.device)
    out = _scan_fwd_raw(u, delta, _f32(A), B, C, _f32(D), z, _f32(delta_bias), delta_softplus, None, 16,
                        h_init=None if h_init is None else h_init.contiguous(), h_last=h_last,
                        variant=SCAN_FWD_VARIANT)
    return out, h_last


# ------------------------------------------------------------------------------------------------
# causal depthwise conv1d + SiLU
# ------------------------------------------------------------------------------------------------
def _conv_args(x, w2, bias, K):
    a = ConvArgs()
    a.struct_size = ct.sizeof(ConvArgs)
    a.dtype = _dtype_code(x)
    a.batch, a.seqlen, a.dim, a.width = x.shape[0], x.shape[1], x.shape[2], K
    a.x, a.x_bs, a.x_ls = _p(x), x.stride(0), x.stride(1)
    a.weight, a.bias = _p(w2), _p(bias)
    return a


class CausalConv1dSiluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, plan=None):
        _require_cuda(x, weight, bias)
        ctx.plan = plan
        D = x.shape[2]
        K = weight.shape[-1]
        if weight.numel() != D * K:
            raise ValueError(f"causal_conv1d: weight {tuple(weight.shape)} must be depthwise [D={D}, 1, K]")
        x = _rows(x)
        w2 = _f32(weight).view(D, K)
        b32 = _f32(bias)
        out = torch.empty((x.shape[0], x.shape[1], D), dtype=x.dtype, device=x.device)
        a = _conv_args(x, w2, b32, K)
        a.out, a.out_bs, a.out_ls = _p(out), out.stride(0), out.stride(1)
        _call("mamba_conv1d_silu_fwd", a, x.device)
        ctx.save_for_backward(x, w2, b32)
        ctx.wshape = weight.shape
        ctx.wdtype = weight.dtype
        ctx.bdtype = None if bias is None else bias.dtype
        return out

    @staticmethod
    def backward(ctx, dout):
        x, w2, b32 = ctx.saved_tensors
        Bsz, L, D = x.shape
        K = w2.shape[1]
        dout = _rows(dout.to(x.dtype))
        plan = ctx.plan
        if plan is not None and plan.xz_shape is not None and plan.xz_shape[-1] == 2 * D:
            dx = plan.part("xz", plan.xz_shape, x.dtype, x.device, 0, D)   # first half of d(xz)
        else:
            dx = torch.empty((Bsz, L, D), dtype=x.dtype, device=x.device)
        dw = torch.empty((D, K), dtype=torch.float32, device=x.device)
        db = torch.empty((D,), dtype=torch.float32, device=x.device) if b32 is not None else None
        wsb = lib().mamba_conv1d_bwd_workspace_bytes(Bsz, L, D, K)
        ws = torch.empty(wsb, dtype=torch.uint8, device=x.device)
        a = _conv_args(x, w2, b32, K)
        a.dout, a.dout_bs, a.dout_ls = _p(dout), dout.stride(0), dout.stride(1)
        a.dx, a.dx_bs, a.dx_ls = _p(dx), dx.stride(0), dx.stride(1)
        a.dweight, a.dbias = _p(dw), _p(db)
        a.workspace, a.workspace_bytes = _p(ws), wsb
        _call("mamba_conv1d_silu_bwd", a, x.device)
        return dx, dw.view(ctx.wshape).to(ctx.wdtype), None if db is None else db.to(ctx.bdtype), None


def causal_conv1d_silu_fn(x, weight, bias=None, plan=None):
    """silu(depthwise causal conv1d(x)) for x [B, L, D] (channels last); weight [D, 1, K] or [D, K]."""
    return CausalConv1dSiluFn.apply(x, weight, bias, plan)


def causal_conv1d_silu_prefill(x, weight, bias=None):
    """Inference-only conv that also returns conv_state [B, D, K] (the last K inputs per channel)."""
    _require_cuda(x, weight, bias)
    x = _rows(x)
    D, K = x.shape[2], weight.shape[-1]
    w2, b32 = _f32(weight).view(D, K), _f32(bias)
    out = torch.empty((x.shape[0], x.shape[1], D), dtype=x.dtype, device=x.device)
    state = torch.empty((x.shape[0], D, K), dtype=x.dtype, device=x.device)
    a = _conv_args(x, w2, b32, K)
    a.out, a.out_bs, a.out_ls = _p(out), out.stride(0), out.stride(1)
    a.final_state = _p(state)
    _call("mamba_conv1d_silu_fwd", a, x.device)
    return out, state


# ------------------------------------------------------------------------------------------------
# RMSNorm (+ residual add)
# ------------------------------------------------------------------------------------------------
def _norm_args(act_dtype, resid_dtype, rows, dim, eps):
    a = NormArgs()
    a.struct_size = ct.sizeof(NormArgs)
    a.dtype = _DTYPES[act_dtype]
    a.resid_dtype = _DTYPES[resid_dtype]
    a.rows, a.dim, a.eps = rows, dim, eps
    return a


class RMSNormFn(torch.autograd.Function):
    """(x [T] | None, weight, residual [TR] | None) -> (y [T], x + residual [TR]).  T = act_dtype."""

    @staticmethod
    def forward(ctx, x, weight, residual, eps, act_dtype):
        _require_cuda(x, weight, residual)
        ref = x if x is not None else residual
        if ref is None:
            raise ValueError("rmsnorm: x and residual are both None")
        shape, dim, dev = ref.shape, ref.shape[-1], ref.device
        T = act_dtype if act_dtype is not None else (x.dtype if x is not None else residual.dtype)
        TR = residual.dtype if residual is not None else T
        if T not in _DTYPES or TR not in _DTYPES:
            raise TypeError(f"rmsnorm: unsupported dtypes ({T}, {TR})")
        x2 = None if x is None else x.to(T).reshape(-1, dim).contiguous()
        r2 = None if residual is None else residual.reshape(-1, dim).contiguous()
        w32 = _f32(weight)
        rows = (x2 if x2 is not None else r2).shape[0]
        y = torch.empty((rows, dim), dtype=T, device=dev)
        # the summed stream is materialised only when there is a sum; otherwise it IS the one operand
        ro = torch.empty((rows, dim), dtype=TR, device=dev) if (x2 is not None and r2 is not None) else None
        rstd = torch.empty((rows,), dtype=torch.float32, device=dev)
        a = _norm_args(T, TR, rows, dim, eps)
        a.x, a.residual, a.weight = _p(x2), _p(r2), _p(w32)
        a.y, a.resid_out, a.rstd = _p(y), _p(ro), _p(rstd)
        _call("mamba_rmsnorm_fwd", a, dev)
        normed = ro if ro is not None else (r2 if r2 is not None else x2)
        ctx.save_for_backward(normed, w32, rstd)
        ctx.meta = (T, TR, eps, shape, weight.dtype, x is not None, residual is not None)
        # with a single operand the stream IS that operand: rmsnorm_fn hands the input itself back so that
        # autograd sees the aliasing (returning a view from here would cut the gradient)
        return y.view(shape), (None if ro is None else ro.view(shape))

    @staticmethod
    def backward(ctx, dy, dres):
        normed, w32, rstd = ctx.saved_tensors
        T, TR, eps, shape, wdtype, has_x, has_res = ctx.meta
        rows, dim = normed.shape
        dev = normed.device
        NT = normed.dtype  # dtype of the tensor that was normalised
        dy2 = dy.to(T).reshape(-1, dim).contiguous()
        both = has_x and has_res
        dr2 = None if (dres is None or not both) else dres.to(TR).reshape(-1, dim).contiguous()
        dx = torch.empty((rows, dim), dtype=T, device=dev) if has_x else None
        dro = torch.empty((rows, dim), dtype=TR, device=dev) if has_res and (not has_x or TR != T) else None
        dw = torch.empty((dim,), dtype=torch.float32, device=dev)
        wsb = lib().mamba_rmsnorm_bwd_workspace_bytes(rows, dim)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        a = _norm_args(T, NT, rows, dim, eps)
        a.residual, a.weight, a.rstd = _p(normed), _p(w32), _p(rstd)
        a.dy, a.dresid_in, a.dweight = _p(dy2), _p(dr2), _p(dw)
        if NT == T:
            a.dx = _p(dx if dx is not None else dro)
            if dx is not None and dro is not None:
                a.dresid_out = _p(dro)
        else:
            a.dx, a.dresid_out = _p(dx), _p(dro)
        a.workspace, a.workspace_bytes = _p(ws), wsb
        _call("mamba_rmsnorm_bwd", a, dev)
        gx = dx.view(shape) if has_x else None
        if has_res:
            gr = (dro if dro is not None else dx).view(shape)
        else:
            gr = None
        return gx, dw.to(wdtype), gr, None, None


def rmsnorm_fn(x, weight, residual=None, eps=1e-5, act_dtype=None):
    """y = rmsnorm(x + residual) * weight.  Returns (y, stream) where stream = x + residual in the
    residual's dtype (or the single operand itself when only one of x / residual is given).
    `act_dtype` is the dtype of y (default: x's, else residual's): the residual stream can stay fp32 while
    the mixer runs in bf16."""
    y, stream = RMSNormFn.apply(x, weight, residual, float(eps), act_dtype)
    if stream is None:
        stream = residual if residual is not None else x
    return y, stream


# ------------------------------------------------------------------------------------------------
# grammar-masked loss (reference train.py:133-138 + :161-165)
# ------------------------------------------------------------------------------------------------
def _loss_args(logits, src, trg, table, boundaries, col_lse, row_lse, ws):
    Bsz, T, V = logits.shape
    a = LossArgs()
    a.struct_size = ct.sizeof(LossArgs)
    a.dtype = _dtype_code(logits)
    a.batch, a.seqlen, a.vocab = Bsz, T, V
    for i in range(4):
        a.boundaries[i] = int(boundaries[i])
    a.logits, a.logits_bs, a.logits_ts = _p(logits), logits.stride(0), logits.stride(1)
    a.src, a.trg, a.table = _p(src), _p(trg), _p(table)
    a.col_lse, a.row_lse = _p(col_lse), _p(row_lse)
    a.workspace, a.workspace_bytes = _p(ws), ws.numel()
    return a


class FilteredCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, src, trg, table, boundaries):
        _require_cuda(logits, src, trg, table)
        if logits.dim() != 3 or logits.stride(2) != 1:
            raise ValueError("filtered_ce: logits must be [B, T, V] with a contiguous vocab axis")
        Bsz, T, V = logits.shape
        if table.shape != (5, V) or table.dtype != torch.float32:
            raise ValueError(f"filtered_ce: table must be fp32 [5, {V}]")
        src = src.reshape(Bsz, T).to(torch.long).contiguous()
        trg = trg.reshape(Bsz, T).to(torch.long).contiguous()
        table = table.contiguous()
        dev = logits.device
        col_lse = torch.empty((Bsz, V), dtype=torch.float32, device=dev)
        row_lse = torch.empty((Bsz, T), dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        ws = torch.empty(lib().mamba_filtered_ce_workspace_bytes(Bsz, T, V), dtype=torch.uint8, device=dev)
        a = _loss_args(logits, src, trg, table, boundaries, col_lse, row_lse, ws)
        a.loss = _p(loss)
        _call("mamba_filtered_ce_fwd", a, dev)
        ctx.save_for_backward(logits, src, trg, table, col_lse, row_lse)
        ctx.boundaries = tuple(int(b) for b in boundaries)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        logits, src, trg, table, col_lse, row_lse = ctx.saved_tensors
        Bsz, T, V = logits.shape
        dev = logits.device
        # same strides as the logits: a padded LM-head output gets a padded gradient (pad columns zeroed once)
        if logits.stride(1) != V or logits.stride(0) != T * V:
            base = torch.zeros((Bsz, T, logits.stride(1)), dtype=logits.dtype, device=dev)
            dlogits = base[:, :, :V]
        else:
            dlogits = torch.empty((Bsz, T, V), dtype=logits.dtype, device=dev)
        ws = torch.empty(lib().mamba_filtered_ce_workspace_bytes(Bsz, T, V), dtype=torch.uint8, device=dev)
        go = grad_out.to(torch.float32).contiguous()
        a = _loss_args(logits, src, trg, table, ctx.boundaries, col_lse, row_lse, ws)
        a.grad_out = _p(go)
        a.dlogits, a.dlogits_bs, a.dlogits_ts = _p(dlogits), dlogits.stride(0), dlogits.stride(1)
        _call("mamba_filtered_ce_bwd", a, dev)
        return dlogits, None, None, None, None


def filtered_ce_fn(logits, src, trg, table, boundaries):
    """mean cross-entropy of the grammar-masked, sequence-normalised logits (see csrc/loss.cu)."""
    return FilteredCEFn.apply(logits, src, trg, table, tuple(boundaries))


# ------------------------------------------------------------------------------------------------
# decode step (no autograd)
# ------------------------------------------------------------------------------------------------
def _step_args(dtype_t, Bsz, D, N, K, R, flags):
    a = StepArgs()
    a.struct_size = ct.sizeof(StepArgs)
    a.dtype = _dtype_code(dtype_t)
    a.batch, a.dim, a.dstate, a.width, a.dt_rank, a.flags = Bsz, D, N, K, R, flags
    return a


@torch.no_grad()
def conv_step(x, conv_state, weight2d, bias):
    """x [B, D] (row stride free), conv_state [B, D, K] updated in place -> xc [B, D]."""
    _require_cuda(x, conv_state, weight2d, bias)
    Bsz, D, K = conv_state.shape
    if x.stride(-1) != 1 or not conv_state.is_contiguous():
        raise ValueError("conv_step: x must have unit inner stride and conv_state must be contiguous")
    xc = torch.empty((Bsz, D), dtype=x.dtype, device=x.device)
    a = _step_args(x, Bsz, D, 0, K, 0, 0)
    a.x, a.x_bs = _p(x), x.stride(0)
    a.conv_state, a.conv_weight, a.conv_bias = _p(conv_state), _p(weight2d), _p(bias)
    a.xc, a.xc_bs = _p(xc), xc.stride(0)
    _call("mamba_conv_step", a, x.device)
    return xc


@torch.no_grad()
def ssm_step(xc, dt_in, Bv, Cv, dt_weight, dt_bias, A, D, z, ssm_state, delta_softplus=True):
    """One recurrence step; ssm_state [B, D, N] fp32 updated in place -> y [B, D]."""
    _require_cuda(xc, dt_in, Bv, Cv, dt_weight, dt_bias, A, D, z, ssm_state)
    Bsz, Dm, N = ssm_state.shape
    R = dt_in.shape[-1]
    for t in (xc, dt_in, Bv, Cv, z):
        if t is not None and t.stride(-1) != 1:
            raise ValueError("ssm_step: activation tensors must have unit inner stride")
    y = torch.empty((Bsz, Dm), dtype=xc.dtype, device=xc.device)
    flags = ((FLAG_HAS_Z if z is not None else 0) | (FLAG_DELTA_SOFTPLUS if delta_softplus else 0)
             | (FLAG_HAS_D if D is not None else 0))
    a = _step_args(xc, Bsz, Dm, N, 0, R, flags)
    a.xc, a.xc_bs = _p(xc), xc.stride(0)
    a.dt_in, a.dt_in_bs = _p(dt_in), dt_in.stride(0)
    a.Bv, a.Bv_bs = _p(Bv), Bv.stride(0)
    a.Cv, a.Cv_bs = _p(Cv), Cv.stride(0)
    a.dt_weight, a.dt_bias, a.A, a.D = _p(dt_weight), _p(dt_bias), _p(A), _p(D)
    if z is not None:
        a.z, a.z_bs = _p(z), z.stride(0)
    a.ssm_state = _p(ssm_state)
    a.y, a.y_bs = _p(y), y.stride(0)
    _call("mamba_ssm_step", a, xc.device)
    return y


@torch.no_grad()
def linear_step(x, weight, bias=None):
    """y = x @ weight.T (+ bias) for x [B <= 16, K]: the weight-streaming decode form of nn.Linear."""
    _require_cuda(x, weight, bias)
    if x.dim() != 2 or x.stride(-1) != 1:
        raise ValueError("linear_step: x must be [B, K] with unit inner stride")
    Bsz, K = x.shape
    N = weight.shape[0]
    if weight.shape[1] != K or not weight.is_contiguous():
        raise ValueError("linear_step: weight must be contiguous [N, K]")
    if bias is not None and bias.dtype != weight.dtype:
        bias = bias.to(weight.dtype)
    y = torch.empty((Bsz, N), dtype=x.dtype, device=x.device)
    a = LinearStepArgs()
    a.struct_size = ct.sizeof(LinearStepArgs)
    a.dtype, a.w_dtype = _dtype_code(x), _dtype_code(weight)
    a.batch, a.in_features, a.out_features = Bsz, K, N
    a.x, a.x_bs = _p(x), x.stride(0)
    a.weight, a.bias = _p(weight), _p(bias)
    a.y, a.y_bs = _p(y), y.stride(0)
    _call("mamba_linear_step", a, x.device)
    return y


def fused_linear_step(x, weight, bias=None, *, norm_weight=None, eps=1e-5, residual_in=None, residual_out=None,
                      conv_state=None, conv_weight=None, conv_bias=None, out=None):
    """linear_step with the residual block's per-token neighbours folded in (decode; include/mamba_b200.h
    mamba_fused_linear_step): optional prologue x := rmsnorm(x + residual_in) * norm_weight (residual_out receives the
    sum; x may be None), optional epilogue on the first conv_state.shape[1] output columns (depthwise conv step + SiLU,
    conv_state updated in place).  Returns y, or (y, conv_out) with an epilogue."""
    _require_cuda(x, weight, bias, norm_weight, residual_in, residual_out, conv_state, conv_weight, conv_bias)
    ref = x if x is not None else residual_in
    Bsz, K = ref.shape
    N = weight.shape[0]
    if weight.shape[1] != K or not weight.is_contiguous():
        raise ValueError("fused_linear_step: weight must be contiguous [N, K]")
    if x is not None and x.stride(-1) != 1:
        raise ValueError("fused_linear_step: x must have unit inner stride")
    if bias is not None and bias.dtype != weight.dtype:
        bias = bias.to(weight.dtype)
    act = x.dtype if x is not None else (conv_state.dtype if conv_state is not None else weight.dtype)
    y = out if out is not None else torch.empty((Bsz, N), dtype=act, device=ref.device)
    a = FusedLinearStepArgs()
    a.struct_size = ct.sizeof(FusedLinearStepArgs)
    a.dtype, a.w_dtype = _DTYPES[act], _dtype_code(weight)
    a.batch, a.in_features, a.out_features = Bsz, K, N
    if x is not None:
        a.x, a.x_bs = _p(x), x.stride(0)
    a.weight, a.bias = _p(weight), _p(bias)
    a.y, a.y_bs = _p(y), y.stride(0)
    if norm_weight is not None:
        if norm_weight.dtype != torch.float32 or (residual_in is not None and residual_in.dtype != torch.float32) or (
                residual_out is not None and residual_out.dtype != torch.float32):
            raise TypeError("fused_linear_step: norm weight and the residual stream are fp32")
        a.norm_weight, a.eps = _p(norm_weight), float(eps)
        if residual_in is not None:
            a.residual_in, a.residual_in_bs = _p(residual_in), residual_in.stride(0)
        if residual_out is not None:
            a.residual_out, a.residual_out_bs = _p(residual_out), residual_out.stride(0)
    conv_out = None
    if conv_state is not None:
        Dc, Kc = conv_state.shape[1], conv_state.shape[2]
        if conv_state.dtype != act or not conv_state.is_contiguous():
            raise ValueError("fused_linear_step: conv_state must be contiguous and of the activation dtype")
        conv_out = torch.empty((Bsz, Dc), dtype=act, device=ref.device)
        a.conv_dim, a.conv_width = Dc, Kc
        a.conv_state, a.conv_weight, a.conv_bias = _p(conv_state), _p(conv_weight), _p(conv_bias)
        a.conv_out, a.conv_out_bs = _p(conv_out), conv_out.stride(0)
    _call("mamba_fused_linear_step", a, ref.device)
    return y if conv_state is None else (y, conv_out)


def sample_step_args(mode, logits, lse, dist, counts, generated, gen_len, next_token, bounds, prompt_len, uniforms=None,
                     win_q=None, win_sum=None):
    """Fill a MambaSampleStepArgs (reused across steps by the decoder; every pointer is a persistent buffer).
    bounds = (dyn, length, time, tempo) first-token ids."""
    _require_cuda(logits, lse, dist, counts, generated, gen_len, next_token, uniforms, win_q, win_sum)
    if logits.dtype != torch.float32 or lse.dtype != torch.float32 or dist.dtype != torch.float32:
        raise TypeError("sample_step: logits, lse and dist are fp32")
    if counts.dtype != torch.int32 or gen_len.dtype != torch.int32 or generated.dtype != torch.int64 or next_token.dtype != torch.int64:
        raise TypeError("sample_step: counts/gen_len int32, generated/next_token int64")
    a = SampleStepArgs()
    a.struct_size = ct.sizeof(SampleStepArgs)
    a.mode, a.batch, a.vocab = int(mode), logits.shape[0], logits.shape[1]
    dyn, length, time, tempo = (int(v) for v in bounds)
    a.bucket_bounds = (ct.c_int32 * 4)(dyn - 1, length - 1, time - 1, tempo - 1)
    a.class_bounds = (ct.c_int32 * 4)(dyn, length, time, tempo)
    if mode == 0:   # scripts/generate_midi_many.py:24-43
        rule, base, cap = (1, 0, 1, 2, 0), (1.04, 1.0, 1.015, 1.0, 1.0), (1.25, 1.0, 1.08, 1.0, 1.0)
    else:           # scripts/generate.py:60-71
        rule, base, cap = (1, 1, 0, 0, 0), (1.01, 1.02, 1.0, 1.0, 1.0), (1.2, 1.2, 1.0, 1.0, 1.0)
    a.pen_rule = (ct.c_int32 * 5)(*rule)
    # min(base ** count, cap) for count 0..127, in python doubles exactly as the reference loops evaluate it (every
    # cap is reached below count 20, so the table's last column is the value for any larger count)
    table = torch.tensor([[min(b ** c, k) for c in range(128)] for b, k in zip(base, cap)], dtype=torch.float64)
    a._pen_table = table.to(torch.float32).to(logits.device).contiguous()    # kept alive by the argument block
    a.pen_table = _p(a._pen_table)
    a.prompt_len, a.time_budget = int(prompt_len), 64 * 16
    a.logits, a.logits_bs = _p(logits), logits.stride(0)
    a.lse, a.dist, a.counts = _p(lse), _p(dist), _p(counts)
    a.generated, a.generated_bs = _p(generated), generated.stride(0)
    a.gen_len, a.next_token = _p(gen_len), _p(next_token)
    a.uniforms, a.win_q, a.win_sum = _p(uniforms), _p(win_q), _p(win_sum)
    return a


def sample_step(args, device):
    _call("mamba_sample_step", args, device)


class DecodeTokenPlan:
    """Argument block of mamba_decode_token (the whole-model, one-launch decode step) for a Layout-P model and its
    inference cache.  Holds every tensor the raw pointers refer to (weights — optionally bf16 copies made here —,
    step constants, the device array of layer descriptors, scratch, barrier words)."""

    SHAPES = ((1024, 2048), (128, 256))

    @staticmethod
    def eligible(model, cache, batch):
        p = getattr(model, "params", None)
        if getattr(model, "layout", None) != "P" or p is None or batch > 16:
            return False
        w = model.layers[0].mixer.in_proj.weight
        return ((p.d_model, p.d_inner) in DecodeTokenPlan.SHAPES and p.d_state % 4 == 0 and p.dt_rank % 4 == 0 and w.is_cuda
                and w.dtype == torch.float32 and all(cs.dtype == torch.float32 for cs, _ in cache)
                and model.lm_head.bias is None and p.d_state <= 64 and p.dt_rank <= 64 and p.d_conv == 4)

    def __init__(self, model, cache, token, logits, weight_dtype=None):
        p = model.params
        dev = token.device
        wdt = weight_dtype or torch.float32
        self.keep = []

        def W(t):   # a weight in the streaming dtype (bf16: a decode-only copy)
            t = t.detach()
            t = t if t.dtype == wdt else t.to(wdt)
            t = t.contiguous()
            self.keep.append(t)
            return t

        def F32(t):
            if t is None:
                return None
            t = t.detach().float().contiguous()
            self.keep.append(t)
            return t

        layers = (DecodeLayer * len(model.layers))()
        for i, (layer, (cs, hs)) in enumerate(zip(model.layers, cache)):
            m, d = layer.mixer, layers[i]
            d.norm_weight = _p(F32(layer.norm.weight))
            d.in_proj_weight, d.in_proj_bias = _p(W(m.in_proj.weight)), _p(None if m.in_proj.bias is None else W(m.in_proj.bias))
            d.conv_weight = _p(F32(m.conv1d.weight).view(p.d_inner, p.d_conv))
            d.conv_bias = _p(F32(m.conv1d.bias))
            d.conv_state, d.ssm_state = _p(cs), _p(hs)
            d.x_proj_weight = _p(W(m.x_proj.weight))
            d.dt_weight, d.dt_bias = _p(F32(m.dt_proj.weight)), _p(F32(m.dt_proj.bias))
            d.A, d.D = _p(F32(-torch.exp(m.A_log.detach().float()))), _p(F32(m.D))
            d.out_proj_weight = _p(W(m.out_proj.weight))
            d.out_proj_bias = _p(None if m.out_proj.bias is None else W(m.out_proj.bias))
        raw = bytes(layers)
        self.layers_dev = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)
        n = lib().mamba_decode_token_scratch_bytes(p.d_model, p.d_inner, p.d_state, p.dt_rank)
        self.scratch = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.barrier = torch.zeros(lib().mamba_decode_token_barrier_bytes(len(model.layers)), dtype=torch.uint8, device=dev)
        self.token, self.logits = token, logits
        a = DecodeTokenArgs()
        a.struct_size = ct.sizeof(DecodeTokenArgs)
        a.w_dtype = _DTYPES[wdt]
        a.batch, a.n_layers, a.vocab = token.shape[0], len(model.layers), model.embedding.weight.shape[0]
        a.d_model, a.d_inner, a.d_state, a.dt_rank, a.d_conv = p.d_model, p.d_inner, p.d_state, p.dt_rank, p.d_conv
        a.eps = float(model.norm_f.eps)
        emb = W(model.embedding.weight)
        a.token, a.embedding, a.layers = _p(token), _p(emb), _p(self.layers_dev)
        a.norm_f_weight, a.head_weight, a.head_bias = _p(F32(model.norm_f.weight)), _p(emb), None   # tied head
        a.logits, a.logits_bs = _p(logits), logits.stride(0)
        a.scratch, a.scratch_bytes, a.barrier = _p(self.scratch), n, _p(self.barrier)
        self.args = a
        self.device = dev

    def run(self):
        _call("mamba_decode_token", self.args, self.device)


def launch_count() -> int:
    return _lib.launch_count()
