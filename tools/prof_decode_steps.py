#!/usr/bin/env python
"""A few eager (un-graphed) decode steps of BASELINE config 4 (10 sequences, default model) for ncu captures."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from mamba_b200 import generate, synthetic, train  # noqa: E402
dev = torch.device("cuda")
torch.manual_seed(0)
model = train.new_model("mamba").to(dev).eval()
src, _, meta = synthetic.batch(10, 512, seed=3)
with torch.no_grad():
    dec = generate.RecurrentDecoder(model, 10, use_graph=False, max_new_tokens=64)
    dec.prefill(src.to(dev), meta.to(dev))
    for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
        dec.step()
torch.cuda.synchronize()
