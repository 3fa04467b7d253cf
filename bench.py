#!/usr/bin/env python
"""bench.py — the reference's headline metric on the B200-native Mamba hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Metric (BASELINE.json): train tokens/s.  Workload at every N: BASELINE configs[1]/[2] — the repo's Mamba
model (Layout P: d_model 1024, 10 layers, d_state 64, expand 2, d_conv 4, vocab 17914, 6 metadata tokens) in
bf16 autocast with an fp32 residual stream and fp32 scan state, per-GPU batch 2 x 2048 synthetic MIDI tokens
(weak scaling, train_parallel.py semantics), one step = forward + grammar-masked loss + backward + gradient
all-reduce (N>1) + Adam, replayed as one CUDA graph per rank.

One JSON line on stdout (rank 0).  `value` is timed with the batch already resident in HBM; `e2e` is the same
step driven from pinned HOST buffers with the loss read back every step.  `roofline` is the dominant kernel
(selective-scan backward) timed live with CUDA events; `cpu_baseline` is the CPU oracle (a port of the
reference's pure-PyTorch Mamba) timed on this box's host cores on a bounded sample.  `--impl reference` times
that CPU implementation alone.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "train_tokens_per_sec"
UNIT = "tokens/s"
WORKLOAD = ("mamba1_layoutP_d1024_L10_N64_vocab17914 train step (fwd+loss+bwd+allreduce+Adam), "
            "per-GPU batch 2 x 2048 tokens (+6 metadata), bf16 autocast / fp32 residual+scan state")


L2_NOTE = ("no explicit flush: one step streams ~1.4 GB of weights/grads/Adam state plus >2 GB of activations, "
           "far above the 126 MB L2")


def workload_config(n_gpus, tokens_per_gpu=4096):
    """`config` of the JSON line — identical for the b200 arm and the reference (CPU) arm at the same N."""
    return {"workload": WORKLOAD, "global_batch_tokens": tokens_per_gpu * n_gpus, "parallelism": f"dp{n_gpus}", "l2": L2_NOTE}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------------------------------
# clocks: sampled DURING the timed region
# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting",
               0x10: "sync_boost"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            log(f"[bench] NVML unavailable ({e}); clocks not sampled")
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(
                    nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._t is not None:
            self._t.join()

    def summary(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def visible_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            pass
    return local_rank


# ----------------------------------------------------------------------------------------------------
# CPU reference arm (oracle): a bounded sample of the same workload on the host cores
# ----------------------------------------------------------------------------------------------------
def cpu_reference_sample(steps, warmup, sample_tokens=None):
    """Full-size model (all 10 layers, every weight shape of the workload) on a SHORT synthetic batch:
    B=1, T=`sample_tokens` (+6 metadata) — per-token cost of the scan and of every GEMM is independent of
    T, so tokens/s of the sample is the CPU path's tokens/s on the workload (the full 2 x 2048 batch needs
    ~180 s and ~18 GB per layer on CPU).  Returns (tokens_per_s, ms_per_step, cores, sample)."""
    import torch
    from mamba_b200 import synthetic
    from oracle import simple_mamba as om
    from oracle import train_ref
    T = int(os.environ.get("MAMBA_B200_CPU_SAMPLE_TOKENS", "506")) if sample_tokens is None else sample_tokens
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    args = om.ModelArgs(d_model=1024, n_layer=10, vocab_size=17914, d_state=64, expand=2, d_conv=4,
                        pad_vocab_size_multiple=1, metadata_vocab_size=568)
    model = om.Mamba(args, scan_impl="unbind")
    opt = torch.optim.Adam(model.parameters(), lr=5e-5)
    times = []
    for i in range(warmup + steps):
        src, trg, meta = synthetic.batch(1, T, seed=100 + i)
        t0 = time.perf_counter()
        out = model(src, meta)
        loss = train_ref.loss_fn(src, trg, out)
        opt.zero_grad()
        loss.backward()
        opt.step()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        log(f"[bench][cpu] step {i} {dt:.2f}s loss {loss.item():.4f}")
    mean = sum(times) / len(times)
    sample = (f"oracle (pure-PyTorch Mamba-1 port, unbind scan) full 10-layer model fwd+loss+bwd+Adam on B=1 x T={T} "
              f"tokens (+6 meta), fp32, {cores} threads, mean of {len(times)} steps")
    return T / mean, mean * 1e3, cores, sample


def cpu_kernel_samples():
    """BASELINE.md section 2 items (i) and (ii) on the host cores, bounded: (i) one MambaBlock forward and
    forward+backward at the repo's layer shape (d_model 1024, d_state 64) on B=1 x T=512; (ii) the kernel-only
    selective_scan of config 5 (L=8192, d_state 16, fp32) on a 128-channel slice of the 2048, forward and
    forward+backward, in the same algorithmic-bytes units as the GPU sweep (cost is linear in B*D)."""
    import torch
    from oracle import simple_mamba as om
    out = {}
    torch.manual_seed(0)
    blk = om.MambaBlock(om.ModelArgs(d_model=1024, n_layer=1, vocab_size=17914, d_state=64, pad_vocab_size_multiple=1))
    x = torch.randn(1, 512, 1024, requires_grad=True)
    t0 = time.perf_counter()
    with torch.no_grad():
        blk(x)
    t_f = time.perf_counter() - t0
    t0 = time.perf_counter()
    blk(x).square().sum().backward()
    t_fb = time.perf_counter() - t0
    out["mamba_block_repo_shape"] = {"sample": "B=1 x T=512, d_model 1024, d_state 64, fp32", "fwd_ms": t_f * 1e3,
                                     "fwd_bwd_ms": t_fb * 1e3, "fwd_tokens_per_s": 512 / t_f, "fwd_bwd_tokens_per_s": 512 / t_fb}
    B, L, D, N = 1, 8192, 128, 16
    g = torch.Generator().manual_seed(0)
    u, dl = torch.randn(B, L, D, generator=g), torch.nn.functional.softplus(torch.randn(B, L, D, generator=g) - 4)
    A = -torch.arange(1, N + 1, dtype=torch.float32).repeat(D, 1)
    Bm, Cm, Dv = torch.randn(B, L, N, generator=g), torch.randn(B, L, N, generator=g), torch.ones(D)
    t0 = time.perf_counter()
    with torch.no_grad():
        om.selective_scan(u, dl, A, Bm, Cm, Dv)
    t_f = time.perf_counter() - t0
    leaves = [t.requires_grad_(True) for t in (u, dl, Bm, Cm)]
    t0 = time.perf_counter()
    om.selective_scan(leaves[0], leaves[1], A, leaves[2], leaves[3], Dv).sum().backward()
    t_fb = time.perf_counter() - t0
    by = algorithmic_bytes(B, L, D, N, 4, 4)
    out["selective_scan_config5"] = {
        "sample": f"B={B}, L={L}, D={D} of 2048 channels, d_state {N}, fp32, oracle selective_scan (python loop over L)",
        "fwd_ms": t_f * 1e3, "fwd_bwd_ms": t_fb * 1e3, "fwd_algo_gbs": by["mamba_scan_fwd"] / t_f / 1e9,
        "fwd_bwd_algo_gbs": (by["mamba_scan_fwd"] + by["mamba_scan_bwd"]) / t_fb / 1e9,
        "fwd_elements_per_s": B * L * D * N / t_f}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # a step of the sample is ~5 s on 16 cores: the driver's usual --steps 20 --warmup 5 runs in full (~2 min)
    steps = max(1, min(args.steps, int(os.environ.get("MAMBA_B200_CPU_MAX_STEPS", "20"))))
    warmup = max(1, min(args.warmup, int(os.environ.get("MAMBA_B200_CPU_MAX_WARMUP", "5"))))
    tps, ms, cores, sample = cpu_reference_sample(steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": tps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": tps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": tps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": ("the reference's own pure-PyTorch Mamba exists only as 3.11 bytecode and its shipped model needs "
                 "mamba_ssm (absent): this arm times the oracle port of that pure-PyTorch path on the host cores; "
                 f"steps clamped to {steps} timed + {warmup} warm-up to stay within minutes"),
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------------
# roofline leg: per-op device times of one eager (un-graphed) step, CUDA events on the launching stream
# ----------------------------------------------------------------------------------------------------
def algorithmic_bytes(B, L, D, N, K, s):
    """bytes that must cross HBM per launch (s = activation element size)."""
    BLD, BLN = B * L * D, B * L * N
    return {
        "mamba_scan_fwd": s * (4 * BLD + 2 * BLN) + 4 * (D * N + 2 * D),
        "mamba_scan_bwd": s * (8 * BLD + 4 * BLN) + 8 * (D * N + 2 * D),
        "mamba_conv1d_silu_fwd": s * 2 * BLD + 4 * (D * K + D),
        "mamba_conv1d_silu_bwd": s * 3 * BLD + 8 * (D * K + D),
    }


MUFU_EX2_PEAK = 4.6e12  # ex2/s, measured on this pool's B200 with tools/microbench.cu (16 lanes/clk/SM at 1.965 GHz)


# FMA-pipe cost of one (b, t, d, n) element in SMSP-cycles per warp-pair of 64 elements, from tools/microbench3.cu:
# two-operand packed FMUL2 2.13 clk, packed FFMA2 with three distinct register pairs 4.27 clk (half lane rate).
# forward: 2 FMUL2 + 2 FFMA2 per state pair; backward (recompute + sweep): 6 FMUL2 + 6 FFMA2.
FMA_CLK_PER_PAIR = {"mamba_scan_fwd": 2 * 2.13 + 2 * 4.27, "mamba_scan_bwd": 6 * 2.13 + 6 * 4.27}
SM_CLOCK_HZ, N_SM = 1.965e9, 148


def fma_view(name, B, L, D, N, mean_us):
    """The scans' other binding pipe: packed fp32 multiply-adds with register operands.  Peak elements/s if the FMA
    pipe of every SM did nothing else (measured instruction rates above)."""
    clk = FMA_CLK_PER_PAIR.get(name)
    if clk is None:
        return None
    peak = 4 * 64 / clk * N_SM * SM_CLOCK_HZ       # 4 SMSPs x 64 elements per warp-pair
    rate = B * L * D * N / (mean_us * 1e-6)
    return {"pipe": "fma_fp32x2_register_operands", "achieved": rate, "peak": peak, "unit": "elements/s", "frac": rate / peak,
            "peak_source": "measured instruction rates (tools/microbench3.cu, profiles/r02_microbench_fma_operands.txt)"}


def mufu_view(name, B, L, D, N, mean_us):
    """The scan kernels execute one (forward) or two (backward) MUFU.EX2 per (b, t, d, n): at d_state 64 that pipe,
    not HBM, is what binds them.  Reported next to the mandatory HBM roofline."""
    per = {"mamba_scan_fwd": 1, "mamba_scan_bwd": 2}.get(name)
    if per is None:
        return None
    rate = per * B * L * D * N / (mean_us * 1e-6)
    return {"pipe": "mufu_ex2", "achieved": rate, "peak": MUFU_EX2_PEAK, "unit": "ex2/s", "frac": rate / MUFU_EX2_PEAK,
            "peak_source": "measured (tools/microbench.cu, gpurun_out/microbench.txt)"}


def kernel_times(trainer, batches, reps=3):
    import torch
    from mamba_b200 import ops
    out = {}
    ops.KERNEL_TIMES = {}
    world, trainer.world_size = trainer.world_size, 1  # rank-local measurement: no collective
    try:
        for r in range(reps):
            src, trg, meta = batches[r % len(batches)]
            trainer.src.copy_(src), trainer.trg.copy_(trg), trainer.meta.copy_(meta)
            torch.cuda._sleep(int(4e6))  # let the host run ahead so the events bracket back-to-back launches
            trainer._step_body()
            torch.cuda.synchronize()
        for name, pairs in ops.KERNEL_TIMES.items():
            ms = [a.elapsed_time(b) for a, b in pairs]
            ms = ms[len(ms) // reps:] if reps > 1 else ms  # drop the first (cold) step
            out[name] = {"launches_per_step": len(pairs) // reps, "mean_us": 1e3 * sum(ms) / len(ms),
                         "min_us": 1e3 * min(ms)}
    finally:
        ops.KERNEL_TIMES = None
        trainer.world_size = world
    return out


def extras(dev, peak):
    """Kernel-only selective-scan sweep at BASELINE configs[4] (L=8192, N=16, D=2048) and at the repo's training
    shape, fp32, CUDA events, L2 flushed between launches."""
    import torch
    from mamba_b200 import ops
    out = {"scan_sweep": [], "peak_gbs": peak}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name, (B, L, D, N) in (("config5_B2_L8192_N16", (2, 8192, 2048, 16)), ("config5_B8_L8192_N16", (8, 8192, 2048, 16)),
                               ("repo_B2_L2054_N64", (2, 2054, 2048, 64))):
        g = torch.Generator(device=dev).manual_seed(0)
        u, z = (torch.randn(B, L, D, device=dev, generator=g) for _ in range(2))
        dl = torch.randn(B, L, D, device=dev, generator=g) - 4
        A = -torch.arange(1, N + 1, device=dev, dtype=torch.float32).repeat(D, 1)
        Bm, Cm = (torch.randn(B, L, N, device=dev, generator=g) for _ in range(2))
        Dv, bias = torch.ones(D, device=dev), torch.zeros(D, device=dev)
        dout = torch.randn(B, L, D, device=dev, generator=g)
        leaves = [t.requires_grad_(True) for t in (u, dl, A, Bm, Cm, Dv, z, bias)]

        def timeit(fn, iters=5):
            fn()
            ts = []
            for _ in range(iters):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            return sorted(ts)[len(ts) // 2]

        with torch.no_grad():
            t_f = timeit(lambda: ops.selective_scan_fn(u, dl, A, Bm, Cm, Dv, z=z, delta_bias=bias, delta_softplus=True))
        o = ops.selective_scan_fn(*leaves[:6], z=leaves[6], delta_bias=leaves[7], delta_softplus=True)
        t_b = timeit(lambda: torch.autograd.grad(o, leaves, dout, retain_graph=True))
        by = algorithmic_bytes(B, L, D, N, 4, 4)
        out["scan_sweep"].append({"shape": name, "dtype": "f32", "fwd_us": t_f, "bwd_us": t_b,
                                  "fwd_gbs": by["mamba_scan_fwd"] / t_f / 1e3, "bwd_gbs": by["mamba_scan_bwd"] / t_b / 1e3,
                                  "fwd_frac_hbm": by["mamba_scan_fwd"] / t_f / 1e3 / peak,
                                  "bwd_frac_hbm": by["mamba_scan_bwd"] / t_b / 1e3 / peak,
                                  "elem_per_s": B * L * D * N / (t_f * 1e-6)})
        del leaves, o, u, z, dl, Bm, Cm, dout
    del flush
    return out


def decode_leg(dev, rank, world, barrier, n_new=2000, n_seq=10, prompt=2048):
    """Autoregressive generation (BASELINE configs[3]): prefill + `n_new` recurrent steps per sequence, greedy
    (scripts/generate_midi_many.py) with the on-device sampler; the sequences are sharded over the ranks."""
    import torch
    import torch.distributed as dist
    from mamba_b200 import generate, synthetic, train
    lo, hi = train.shard_rows(n_seq, rank, world)
    torch.manual_seed(0)
    model = train.new_model("mamba").to(dev).eval()
    src, _, meta = synthetic.batch(n_seq, prompt, seed=3)
    ms, dec = 0.0, None
    if hi > lo:
        with torch.no_grad():
            dec = generate.RecurrentDecoder(model, hi - lo, use_graph=True, max_new_tokens=n_new + 16)
            dec.prefill(src[lo:hi].to(dev), meta[lo:hi].to(dev))
            for _ in range(8):
                dec.step()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n_new - 9):
                dec.step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            toks = dec.tokens()
    else:
        barrier()
    barrier()
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    n = n_new - 9
    return {"sequences": n_seq, "sequences_per_rank_max": -(-n_seq // world), "prompt": prompt, "new_tokens_per_seq": n_new,
            "timed_steps": n, "dtype": "f32", "ms_per_step": ms / n, "tokens_per_sec": n_seq * n / (ms * 1e-3),
            "sharding": f"{n_seq} sequences over {world} rank(s), no collective",
            "mode": ("one persistent cooperative kernel per token (csrc/decode.cu: all layers + head, grid barrier between phases) + "
                     "the on-device sampler, one CUDA graph per token, greedy" if getattr(dec, "plan", None) is not None else
                     "recurrent decode: 4 launches per layer + head + on-device sampler, one CUDA graph per token, greedy")}


# ----------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # the scans run one CTA per SM on 128 of the 148 SMs; NCCL's all-reduce kernels share the GPU with them for most
        # of the backward.  Capping NCCL at 16 CTAs keeps them inside the 20 SMs the scans leave free
        # (measured at 8 GPUs: 13.34 ms/step against 13.55-13.64 with NCCL's default, 13.65 with 8 CTAs)
        os.environ.setdefault("NCCL_MAX_CTAS", "16")
        dist.init_process_group("nccl", device_id=dev)
    if world != args.gpus and rank == 0:
        log(f"[bench] WORLD_SIZE={world} but --gpus {args.gpus}: using WORLD_SIZE")

    from mamba_b200 import _lib, synthetic, train
    from mamba_b200.configs import common as cc
    _lib.lib()  # fail loudly if the CUDA extension is missing

    B, T = cc.config.values.batch_size, cc.config.values.block_len
    torch.manual_seed(0)
    model = train.new_model("mamba", layout="P").to(dev)
    adt = torch.bfloat16 if args.dtype == "bf16" else None
    trainer = train.Trainer(model, autocast_dtype=adt, world_size=world, use_graph=not args.no_graph,
                            overlap_allreduce=os.environ.get("MAMBA_B200_OVERLAP", "0") == "1")

    nb = 8
    host = [tuple(t.pin_memory() for t in synthetic.batch(B, T, seed=1000 * rank + i)) for i in range(nb)]
    resident = [tuple(t.to(dev) for t in b) for b in host]

    # launches of OUR kernels per step, counted on an eager step (graph replays do not pass through the host)
    n0 = _lib.launch_count()
    trainer.src.copy_(resident[0][0]), trainer.trg.copy_(resident[0][1]), trainer.meta.copy_(resident[0][2])
    saved = [p.detach().clone() for p in model.parameters()]
    trainer._step_body()
    torch.cuda.synchronize()
    launches_per_step = _lib.launch_count() - n0
    with torch.no_grad():
        for p, q in zip(model.parameters(), saved):
            p.copy_(q)
        for st in trainer.optimizer.state.values():
            for v in st.values():
                if torch.is_tensor(v):
                    v.zero_()
    del saved

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(batches, read_loss):
        for i in range(args.warmup):
            loss = trainer.step(*batches[i % nb])
            if read_loss:
                loss.item()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(visible_index(local)) as cs:
            e0.record()
            for i in range(args.steps):
                loss = trainer.step(*batches[i % nb])
                if read_loss:
                    last = loss.item()
            e1.record()
            barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, cs.summary(), float(trainer.loss.item())

    ms_dev, clocks, loss_dev = timed(resident, read_loss=False)
    ms_e2e, clocks_e2e, loss_e2e = timed(host, read_loss=True)
    tokens = B * T * world * args.steps
    value = tokens / (ms_dev * 1e-3)
    e2e = tokens / (ms_e2e * 1e-3)
    h2d = sum(t.numel() * t.element_size() for t in host[0])

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic",
        "config": workload_config(world, B * T),
        "cuda_graph": not args.no_graph,
        "clocks": clocks,
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps, "clocks": clocks_e2e},
        "gpu_launches": launches_per_step * args.steps,
        "gpu_launches_per_step": launches_per_step,
        "loss": loss_dev,
    }

    # ---- BASELINE config 4: 2000 new tokens for 10 sequences (5 composer bands x 2), sharded by sequence over the ranks
    #      (no collective on the data path; the barrier and the MAX over ranks are measurement only) ---------------------
    if not args.no_extras:
        try:
            line["decode"] = decode_leg(dev, rank, world, barrier)
        except Exception as e:
            line["decode"] = {"error": str(e)[:200]}
    if world > 1:
        # data-parallel correctness in the record: every rank must hold the same parameters after the timed steps
        flat = torch.cat([p.detach().flatten().float() for p in model.parameters()])
        hi, lo = flat.clone(), flat.clone()
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        line["cross_rank_param_max_abs_diff"] = float((hi - lo).max())
        line["param_checksum_rank0"] = float(flat.double().sum())
        del flat, hi, lo
    if world > 1:
        # Every collective of the run is behind us.  Ranks > 0 leave now; rank 0 goes on with rank-local
        # measurements.  (No destroy_process_group: tearing NCCL down under a live CUDA graph that holds captured
        # all-reduces can hang; the processes exit instead.)
        dist.barrier()
        torch.cuda.synchronize()
        if rank != 0:
            sys.stdout.flush()
            os._exit(0)
    if rank == 0:
        # ---- roofline of the dominant kernel, timed live ------------------------------------------------
        try:
            peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
            peak, peak_src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        kt = kernel_times(trainer, resident)
        p = model.params
        ab = algorithmic_bytes(B, T + cc.N_META, p.d_inner, p.d_state, p.d_conv, 2 if args.dtype == "bf16" else 4)
        per_kernel = {}
        for name, st in kt.items():
            e = dict(st)
            if name in ab:
                e["algorithmic_bytes"] = ab[name]
                e["achieved_gbs"] = ab[name] / (st["mean_us"] * 1e-6) / 1e9
                e["frac_of_hbm_peak"] = e["achieved_gbs"] / peak
            e["share_of_step"] = st["mean_us"] * st["launches_per_step"] / (1e3 * ms_dev / args.steps)
            per_kernel[name] = e
        dom = max((n for n in per_kernel if n in ab), key=lambda n: per_kernel[n]["mean_us"] * per_kernel[n]["launches_per_step"])
        d = per_kernel[dom]
        # DRAM traffic of the dominant kernel per launch, from the committed ncu --set full capture of the same
        # shape (only valid for the bf16 default workload it was taken on)
        traffic, traffic_src = None, None
        if args.dtype == "bf16":
            import hashlib
            csrc = ROOT / "deep-learning-based-sequence-models-for-music-generation_b200" / "csrc"
            for tag in ("r02", "r01"):   # newest capture of this kernel at this shape
                try:
                    rec = json.loads((ROOT / "profiles" / f"{tag}_ncu_traffic.json").read_text())[dom]
                    if "source_sha" in rec:   # stale capture (the kernel's sources changed since): report null, not an old number
                        h = hashlib.sha256()
                        for f in rec["source_files"]:
                            h.update((csrc / f).read_bytes())
                        if h.hexdigest()[:16] != rec["source_sha"]:
                            traffic_src = f"profiles/{tag}_ncu_traffic.json is stale (kernel sources changed since the capture)"
                            break
                    traffic, traffic_src = rec["traffic_bytes"], f"profiles/{tag}_ncu_traffic.json ({rec.get('capture', 'ncu --set full')})"
                    break
                except Exception:
                    continue
        line["roofline"] = {"kernel": dom, "bound": "hbm", "achieved": d["achieved_gbs"], "peak": peak, "unit": "GB/s",
                            "frac": d["frac_of_hbm_peak"], "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                            "mean_us": d["mean_us"], "algorithmic_bytes": d["algorithmic_bytes"],
                            "binding_pipe": mufu_view(dom, B, T + cc.N_META, p.d_inner, p.d_state, d["mean_us"]),
                            "binding_pipes": [v for v in (mufu_view(dom, B, T + cc.N_META, p.d_inner, p.d_state, d["mean_us"]),
                                                          fma_view(dom, B, T + cc.N_META, p.d_inner, p.d_state, d["mean_us"])) if v],
                            "kernel_times": "eager step bracketed by CUDA events per C-ABI call (shares of the step; `value` "
                                            "times the graph replay)",
                            "note": "d_state=64 makes this kernel bound by the MUFU and FMA pipes and by issue slots together, "
                                    "not by HBM (SURVEY.md F8, DESIGN.md section 4); the fractions are of each pipe alone on "
                                    "all 148 SMs (the grid covers 128); see profiles/ for the ncu pipe utilisation"}
        line["kernels"] = per_kernel
        # ---- the other two figures of BASELINE.json's metric, N=1 only (short runs; tools/bench_scan.py and
        #      tools/bench_decode.py are the full versions) ---------------------------------------------------------
        if world == 1 and not args.no_extras:
            try:
                line["extras"] = extras(dev, peak)
            except Exception as e:  # never lose the headline line to an extra
                line["extras"] = {"error": str(e)[:200]}
        # ---- CPU baseline (oracle port) on this box's host cores, N=1 only ----------------------------------
        if world == 1 and not args.no_cpu_baseline:
            tps, ms, cores, sample = cpu_reference_sample(steps=2, warmup=1)
            line["cpu_baseline"] = {"value": tps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                                    "ms_per_step": ms}
            try:
                line["cpu_baseline"]["kernels"] = cpu_kernel_samples()
            except Exception as e:
                line["cpu_baseline"]["kernels"] = {"error": str(e)[:200]}
        print(json.dumps(line), flush=True)
    if world > 1:
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    return 0


if __name__ == "__main__":
    sys.exit(main())
