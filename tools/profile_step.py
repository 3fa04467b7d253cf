#!/usr/bin/env python
"""Kernel-time breakdown of one (eager, un-graphed) training step with torch.profiler — a quick look at where the
step goes between the ncu launch lists kept under profiles/."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mamba_b200 import synthetic, train  # noqa: E402
from mamba_b200.configs import common as cc  # noqa: E402

dev = torch.device("cuda")
torch.manual_seed(0)
model = train.new_model("mamba").to(dev)
tr = train.Trainer(model, autocast_dtype=torch.bfloat16, use_graph=False)
b = [t.to(dev) for t in synthetic.batch(cc.config.values.batch_size, cc.config.values.block_len, seed=1)]
for _ in range(3):
    tr.step(*b)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.step(*b)
    torch.cuda.synchronize()
ev = prof.key_averages()
rows = sorted(ev, key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print(f"total device time {tot / 1e3:.2f} ms over {sum(e.count for e in rows)} launches")
for e in rows[:40]:
    print(f"{e.device_time_total / 1e3:8.3f} ms {e.count:4d} {100 * e.device_time_total / tot:5.1f}%  {e.key[:110]}")
