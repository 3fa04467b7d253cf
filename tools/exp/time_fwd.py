"""Development: time the TMA forward from an experiment build of the library (tools/exp/libexp*.so)."""
import json, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from mamba_b200 import _lib
if len(sys.argv) > 1 and sys.argv[1] != "base":
    _lib.LIB_PATH = Path(sys.argv[1]).resolve()
from mamba_b200 import ops
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, (B, L, D, N), dt, variants in [("repo_bf16", (2, 2054, 2048, 64), torch.bfloat16, (100, 110, 102, 112)),
                                         ("long_f32", (2, 8192, 2048, 16), torch.float32, (100, 110))]:
    g = torch.Generator(device="cuda").manual_seed(0)
    r = lambda *s: torch.randn(*s, device="cuda", generator=g)
    u, z, dl = r(B, L, D).to(dt), r(B, L, D).to(dt), (r(B, L, D) - 3).to(dt)
    A = -(torch.rand(D, N, device="cuda", generator=g) * 8 + 0.05)
    Bm, Cm, Dv, bias = r(B, L, N).to(dt), r(B, L, N).to(dt), torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
    out = torch.empty_like(u)
    for v in variants:
        fn = lambda: ops._scan_fwd_raw(u, dl, A, Bm, Cm, Dv, z, bias, True, None, 16, variant=v, out=out)
        for _ in range(2):
            fn()
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        print(json.dumps({"lib": sys.argv[1] if len(sys.argv) > 1 else "base", "shape": name, "variant": v, "median_us": round(ts[2], 1)}), flush=True)
