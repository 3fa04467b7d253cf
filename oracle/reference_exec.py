"""ORACLE tooling (test infrastructure): classes whose methods ARE the reference's bytecode.

`load_reference()` reads `/root/reference/models/mamba/__pycache__/simple_mamba.cpython-311.pyc` and returns
nn.Module classes whose `__init__` / `forward` / `ssm` / `selective_scan` are the reference's own code objects,
executed by the interpreter in `oracle/pyc311.py` against the real torch / einops in this container:

    RMSNorm.__init__, RMSNorm.forward, MambaBlock.__init__, MambaBlock.forward, MambaBlock.ssm,
    MambaBlock.selective_scan, ResidualBlock.__init__, ResidualBlock.forward, Mamba.forward,
    ModelArgs.__post_init__

Only `Mamba.__init__` is not executed (it builds its layer list with a closure-carrying list comprehension);
its six assignments are restated in `RefMamba.__init__` below and the layers it creates are the executed
`ResidualBlock`s.  Used by tests/golden/make_golden.py (fixture generation) and, when /root/reference is present,
by tests/test_oracle.py for a live bit-for-bit comparison with the restatement in oracle/simple_mamba.py.
"""
from __future__ import annotations

import math
from pathlib import Path
from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F
from einops import einsum, rearrange, repeat

from . import pyc311

REF_PYC = Path("/root/reference/models/mamba/__pycache__/simple_mamba.cpython-311.pyc")


def _super_proxy(obj):
    """Stand-in for zero-argument super() inside an nn.Module subclass: the bytecode only calls __init__ on it."""
    return SimpleNamespace(__init__=lambda *a, **k: nn.Module.__init__(obj, *a, **k))


def available() -> bool:
    return REF_PYC.exists()


def load_reference(path: Path = REF_PYC) -> SimpleNamespace:
    root = pyc311.load_pyc(path)
    ns = SimpleNamespace()
    globs = dict(torch=torch, nn=nn, F=F, math=math, einsum=einsum, rearrange=rearrange, repeat=repeat)

    def fn(qualname, defaults=()):
        return pyc311.Function(pyc311.find_code(root, qualname), globs, defaults)

    def init_with_super(qualname, defaults=()):
        code = pyc311.find_code(root, qualname)

        def __init__(self, *args, **kwargs):
            g = dict(globs)
            g["super"] = lambda: _super_proxy(self)   # the bytecode calls super().__init__()
            return pyc311.run(code, g, (self,) + args, kwargs, defaults)
        return __init__

    class RMSNorm(nn.Module):
        __init__ = init_with_super("RMSNorm.__init__", defaults=(1e-5,))
        forward = fn("RMSNorm.forward")

    class MambaBlock(nn.Module):
        __init__ = init_with_super("MambaBlock.__init__")
        forward = fn("MambaBlock.forward")
        ssm = fn("MambaBlock.ssm")
        selective_scan = fn("MambaBlock.selective_scan")

    class ResidualBlock(nn.Module):
        __init__ = init_with_super("ResidualBlock.__init__")
        forward = fn("ResidualBlock.forward")

    class Mamba(nn.Module):
        def __init__(self, params):  # restated (simple_mamba @L58-70); see module docstring
            super().__init__()
            self.params = params
            self.vocab_size = params.vocab_size
            self.metadata_vocab_size = params.metadata_vocab_size
            self.embedding = nn.Embedding(params.vocab_size, params.d_model)
            self.metadata_embedding = nn.Embedding(params.metadata_vocab_size, params.d_model)
            self.layers = nn.ModuleList([ResidualBlock(params) for _ in range(params.n_layer)])
            self.norm_f = RMSNorm(params.d_model)
            self.lm_head = nn.Linear(params.d_model, params.vocab_size, bias=False)
            self.lm_head.weight = self.embedding.weight

        forward = fn("Mamba.forward")

    globs.update(RMSNorm=RMSNorm, MambaBlock=MambaBlock, ResidualBlock=ResidualBlock, Mamba=Mamba)
    ns.RMSNorm, ns.MambaBlock, ns.ResidualBlock, ns.Mamba = RMSNorm, MambaBlock, ResidualBlock, Mamba
    ns.post_init = fn("ModelArgs.__post_init__")
    ns.root = root
    return ns


def make_params(ref: SimpleNamespace, **kw) -> SimpleNamespace:
    """A params object as the reference builds it: dataclass defaults (simple_mamba @L33-44) + the executed
    `ModelArgs.__post_init__` (@L46-54)."""
    base = dict(d_state=16, expand=2, dt_rank="auto", d_conv=4, pad_vocab_size_multiple=8, conv_bias=True, bias=False,
                metadata_vocab_size=568)
    base.update(kw)
    p = SimpleNamespace(**base)
    ref.post_init(p)
    return p
