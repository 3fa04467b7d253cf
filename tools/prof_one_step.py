#!/usr/bin/env python
"""One eager (un-graphed) training step of the benchmarked configuration, for ncu captures of the step's kernels."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mamba_b200 import synthetic, train  # noqa: E402
from mamba_b200.configs import common as cc  # noqa: E402
dev = torch.device("cuda")
torch.manual_seed(0)
model = train.new_model("mamba", layout="P").to(dev)
tr = train.Trainer(model, autocast_dtype=torch.bfloat16, use_graph=False)
b = [t.to(dev) for t in synthetic.batch(cc.config.values.batch_size, cc.config.values.block_len, seed=1)]
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    tr.step(*b)
torch.cuda.synchronize()
