"""Synthetic composer-conditioned MIDI-token batches; there is no dataset on the box.

Tokens follow the reference grammar so that the grammar-masked loss (train.filtered_logit) is exercised:
position p carries class p mod 5 in (pitch, dyn, length, time, tempo) (reference processing/processing.py
encode :129-152 emits 5 tokens per note) with a value uniform in that class's id range
(configs/common/__init__.py:42-57).  Metadata rows are [band, genre x4, decade] (reference
processing/dataset.py:126-130) drawn from the id ranges of the reference tokenization.json.
"""
from __future__ import annotations

import torch

from .configs import common as cc

_CLASSES = ("pitch", "dyn", "length", "time", "tempo")


def token_sequences(batch: int, length: int, seed: int = 0) -> torch.Tensor:
    """[batch, length] int64 on the CPU, deterministic in `seed`."""
    g = torch.Generator().manual_seed(seed)
    starts = [cc.start_idx[c] for c in _CLASSES] + [cc.vocab_size]
    lo = torch.tensor([starts[i] for i in range(5)])
    hi = torch.tensor([starts[i + 1] for i in range(5)])
    cls = torch.arange(length) % 5
    u = torch.rand(batch, length, generator=g)
    return (lo[cls] + (u * (hi[cls] - lo[cls]).float()).long()).clamp_(max=cc.vocab_size - 1)


def metadata(batch: int, seed: int = 0) -> torch.Tensor:
    """[batch, 6] int64 = [band, genre, genre, genre, genre, decade]."""
    g = torch.Generator().manual_seed(seed + 7919)
    r = cc.meta_ranges

    def draw(lo, hi, n):
        return torch.randint(lo, hi + 1, (batch, n), generator=g)

    return torch.cat((draw(*r["band"], 1), draw(*r["genre"], 4), draw(*r["decade"], 1)), dim=1)


def batch(batch_size: int = None, block_len: int = None, seed: int = 0):
    """(src [B, T], trg [B, T], meta [B, 6]) as the reference SequenceDataset.__getitem__ returns them
    (processing/dataset.py:195: sequence[:-1], sequence[1:], band_metadata)."""
    B = cc.config.values.batch_size if batch_size is None else batch_size
    T = cc.config.values.block_len if block_len is None else block_len
    seq = token_sequences(B, T + 1, seed)
    return seq[:, :-1].contiguous(), seq[:, 1:].contiguous(), metadata(B, seed)
