"""Development: per-phase time stamps of the persistent decode kernel (MAMBA_DECODE_FLAG_STAMPS)."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from mamba_b200 import generate, synthetic, train
dev = torch.device("cuda")
torch.manual_seed(0)
nseq = int(sys.argv[1]) if len(sys.argv) > 1 else 10
model = train.new_model("mamba").to(dev).eval()
src, _, meta = synthetic.batch(nseq, 256, seed=3)
NAMES = {1: "P1 end", 2: "bar", 3: "P2 end", 4: "bar", 5: "P3 end", 6: "bar", 7: "P4 end", 8: "bar", 9: "head end",
         10: "P1 staged", 11: "P1 mul done", 12: "P3 staged", 13: "P3 dt done", 14: "P4 mul done"}
with torch.no_grad():
    dec = generate.RecurrentDecoder(model, nseq, use_graph=False, max_new_tokens=64)
    dec.prefill(src.to(dev), meta.to(dev))
    dec.plan.args.flags = 1
    for _ in range(5):
        dec.plan.barrier.zero_()
        dec.step()
    torch.cuda.synchronize()
    st = dec.plan.barrier[16:].view(torch.int64).cpu().tolist()
ev = [(v >> 56, v & ((1 << 56) - 1)) for v in st if v != 0]
print("total us", (ev[-1][1] - ev[0][1]) / 1e3, "events", len(ev))
per = len([1 for i, _ in ev if i not in (0, 9)]) // 10
tot = {}
for k in range(1, len(ev)):
    i, t = ev[k]
    layer = (k - 1) // per
    dt = (t - ev[k - 1][1]) / 1e3
    tot[(i, NAMES.get(i, str(i)))] = tot.get((i, NAMES.get(i, str(i))), 0) + dt
    if layer == 5 or i == 9:
        print(f"  layer {layer} -> {NAMES.get(i, i)}: {dt:.2f}")
print({f"{k[1]}#{k[0]}": round(v, 1) for k, v in tot.items()})
