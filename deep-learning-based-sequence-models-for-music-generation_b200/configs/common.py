"""Mirror of `configs.common` (reference configs/common/__init__.py:19-57, config.yaml:1-27).

Same public names: `config` (SimpleNamespace tree), `vocab_size`, `metadata_vocab_size`, `start_idx`.
`metadata_vocab_size` is the VOCAB_SIZE entry of the reference's tokenization.json (:576); the reference
reads it from /scratch at import, here it is a constant (override with MAMBA_B200_META_VOCAB).
`config.values.device` defaults to 'cuda' as in config.yaml:14 (override with MAMBA_B200_DEVICE).
"""
from __future__ import annotations

import os
from types import SimpleNamespace


def dict_to_namespace(d):
    if isinstance(d, dict):
        return SimpleNamespace(**{k: dict_to_namespace(v) for k, v in d.items()})
    return d


config = dict_to_namespace({
    "discretization": {"pitch": 128, "dyn": 128, "length": 512, "time": 512, "channel": 129, "tempo": 250},
    "resolution": {"bar_res": 64},
    "values": {
        "block_len": 2048,
        "device": os.environ.get("MAMBA_B200_DEVICE", "cuda"),
        "metadata_dims": {"composer": 8},
        "dropout": 0.01,
        "epochs": 10000,
        "eval_interval": 10,
        "save_interval": 10,
        "learning_rate": 0.00005,
        "eval_iters": 200,
        "test_ratio": 0.2,
        "batch_size": 2,
        "augmentation": False,
        "end_of_seq": False,
        "start_of_seq": False,
        "parallel": False,
    },
})

_d = config.discretization
vocab_size = sum([_d.pitch * _d.channel, _d.dyn, _d.length, _d.time, _d.tempo])  # 17914

metadata_vocab_size = int(os.environ.get("MAMBA_B200_META_VOCAB", "568"))
N_META = 6  # band + 4 genres + decade (reference processing/dataset.py:126-130)

start_idx = {}
_off = 0
for _name, _size in (("pitch", _d.pitch * _d.channel), ("dyn", _d.dyn), ("length", _d.length), ("time", _d.time),
                     ("tempo", _d.tempo)):
    start_idx[_name] = _off
    _off += _size
del _off, _name, _size

# id ranges of the metadata vocabulary (reference tokenization.json; processing/dataset.py:103-121)
meta_ranges = {"decade": (1, 201), "decade_null": 0, "genre": (203, 311), "genre_null": 202,
               "band": (313, 567), "band_null": 312}
