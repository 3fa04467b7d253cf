import sys, torch
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from mamba_b200 import generate, synthetic, train
from mamba_b200.configs import common as cc
dev = torch.device("cuda")
torch.manual_seed(0)
model = train.new_model("mamba").to(dev).eval()
src, _, meta = synthetic.batch(10, 2048, seed=3)
src, meta = src.to(dev), meta.to(dev)
with torch.no_grad():
    dec = generate.RecurrentDecoder(model, 10, use_graph=False)
    dec.prefill(src, meta)
    for _ in range(5): dec.step()
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(10): dec.step()
        torch.cuda.synchronize()
ev = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in ev)
print(f"total device time per step {tot/1e3/10:.3f} ms over {sum(e.count for e in ev)/10:.0f} launches")
for e in ev[:16]:
    print(f"{e.device_time_total/1e3/10:8.4f} ms/step {e.count/10:5.1f} {100*e.device_time_total/tot:5.1f}%  {e.key[:100]}")
