#!/usr/bin/env python
"""Decode benchmark (BASELINE configs[3]): 5 composer conditions x 2 samples = 10 sequences, 2048-token prompt,
N new tokens each, greedy, recurrent step kernels under one CUDA graph per token.  Batch-sharded across ranks when
launched with torchrun (no collective on this path).  Prints one JSON line per rank-0."""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mamba_b200 import generate, synthetic, train  # noqa: E402
from mamba_b200.configs import common as cc  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=2000)
    ap.add_argument("--seqs", type=int, default=10)
    ap.add_argument("--prompt", type=int, default=cc.config.values.block_len)
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16"])
    ap.add_argument("--literal-steps", type=int, default=3, help="steps of the reference-style full re-forward loop to time")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lo, hi = train.shard_rows(args.seqs, rank, world)
    torch.manual_seed(0)
    model = train.new_model("mamba").to(dev).eval()
    if args.dtype == "bf16":
        model = model.to(torch.bfloat16)
    src, _, meta = synthetic.batch(args.seqs, args.prompt, seed=3)
    src, meta = src[lo:hi].to(dev), meta[lo:hi].to(dev)
    with torch.no_grad():
        dec = generate.RecurrentDecoder(model, hi - lo, use_graph=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        first = dec.prefill(src, meta)
        torch.cuda.synchronize()
        t_prefill = time.perf_counter() - t0
        out = [first]
        for _ in range(8):  # warm-up incl. graph capture
            out.append(dec.step().clone())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = args.tokens - len(out)
        for _ in range(n):
            out.append(dec.step().clone())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        # the reference's own loop shape: full forward over the window per token (on the same kernels)
        t_lit = None
        if args.literal_steps > 0:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            generate.generate_literal(model, args.prompt, src, meta, args.literal_steps)
            torch.cuda.synchronize()
            t_lit = (time.perf_counter() - t0) / args.literal_steps
    if rank == 0:
        line = {"metric": "decode_tokens_per_sec", "seqs_total": args.seqs, "seqs_this_rank": hi - lo, "world": world,
                "new_tokens": args.tokens, "dtype": args.dtype, "prompt": args.prompt,
                "ms_per_step": ms / n, "tokens_per_sec_this_rank": (hi - lo) * n / (ms * 1e-3),
                "prefill_s": t_prefill,
                "literal_full_reforward_s_per_token_step": t_lit,
                "speedup_vs_literal_loop": (t_lit / (ms / n * 1e-3)) if t_lit else None}
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
