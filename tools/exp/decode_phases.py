"""Development: per-phase time stamps of the persistent decode kernel (library built with -DMB_DEC_PROFILE)."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from mamba_b200 import _lib
_lib.LIB_PATH = ROOT / "tools" / "exp" / "libdecprof.so"
from mamba_b200 import generate, ops, synthetic, train
dev = torch.device("cuda")
torch.manual_seed(0)
model = train.new_model("mamba").to(dev).eval()
src, _, meta = synthetic.batch(10, 256, seed=3)
orig = ops.DecodeTokenPlan.__init__
def patched(self, *a, **k):
    orig(self, *a, **k)
    self.barrier = torch.zeros(2 + 2 + 2 * 200, dtype=torch.int32, device=dev)
    self.args.barrier = self.barrier.data_ptr()
ops.DecodeTokenPlan.__init__ = patched
with torch.no_grad():
    dec = generate.RecurrentDecoder(model, 10, use_graph=False, max_new_tokens=64)
    dec.prefill(src.to(dev), meta.to(dev))
    for _ in range(5):
        dec.step()
    torch.cuda.synchronize()
    st = dec.plan.barrier[4:].view(torch.int64).cpu().tolist()
n = 1 + 10 * 8 + 1
st = st[:n]
d = [b - a for a, b in zip(st[:-1], st[1:])]
names = ["P1 in_proj", "bar", "P2 x_proj", "bar", "P3 ssm", "bar", "P4 out_proj", "bar"]
print("total us", (st[-1] - st[0]) / 1e3)
for l in (0, 1, 5, 9):
    print("layer", l, " ".join(f"{names[i]}={d[l * 8 + i] / 1e3:.1f}" for i in range(8)))
print("head us", d[80] / 1e3)
tot = {}
for l in range(10):
    for i in range(8):
        tot[names[i] + str(i)] = tot.get(names[i] + str(i), 0) + d[l * 8 + i] / 1e3
print({k: round(v, 1) for k, v in tot.items()})
