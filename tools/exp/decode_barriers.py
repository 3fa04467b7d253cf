"""Development: per-CTA arrive/leave times of every grid barrier of the persistent decode kernel."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from mamba_b200 import generate, synthetic, train
dev = torch.device("cuda")
torch.manual_seed(0)
model = train.new_model("mamba").to(dev).eval()
src, _, meta = synthetic.batch(10, 256, seed=3)
with torch.no_grad():
    dec = generate.RecurrentDecoder(model, 10, use_graph=False, max_new_tokens=64)
    dec.prefill(src.to(dev), meta.to(dev))
    dec.plan.args.flags = 3
    for _ in range(5):
        dec.step()
    torch.cuda.synchronize()
    L = 10
    st = dec.plan.barrier[16:].view(torch.int64)[16 * L + 8:].cpu()
ncta = 148
t = st[: ncta * 2 * 4 * L].view(ncta, 4 * L, 2).double()
arr, lv = t[:, :, 0], t[:, :, 1]
names = ["after P1", "after P2", "after P3", "after P4"]
skew = (arr.max(0).values - arr.min(0).values) / 1e3
lat = (lv.min(0).values - arr.max(0).values) / 1e3
lat2 = (lv.max(0).values - arr.max(0).values) / 1e3
for k in range(4):
    sel = slice(4 + k, 4 * L, 4)
    print(f"{names[k]}: arrival skew mean {skew[sel].mean():.2f} us (max {skew[sel].max():.2f}); last arrival -> first leave {lat[sel].mean():.2f} us, -> last leave {lat2[sel].mean():.2f} us")
# which CTAs arrive last?
last = arr.argmax(0)
print("last-arriving CTA histogram (top 8):", torch.bincount(last, minlength=ncta).topk(8))
ph = (arr[:, 1:] - lv[:, :-1]) / 1e3   # phase durations per CTA
for k in range(4):
    sel = slice(4 + k, 4 * L - 1, 4)
    d = ph[:, sel]
    print(f"phase after barrier '{names[k]}': per-CTA duration min {d.min():.2f} median {d.median():.2f} max {d.max():.2f} us")
