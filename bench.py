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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps = max(1, min(args.steps, int(os.environ.get("MAMBA_B200_CPU_MAX_STEPS", "6"))))
    warmup = max(1, min(args.warmup, 1))
    tps, ms, cores, sample = cpu_reference_sample(steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": tps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "parallelism": "cpu"},
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
    """(a) kernel-only selective-scan sweep at BASELINE configs[4] (L=8192, N=16, D=2048) and at the repo's training
    shape, fp32, CUDA events, L2 flushed between launches; (b) recurrent greedy decode, 10 sequences (5 composer
    conditions x 2), 2048-token prompt."""
    import torch
    from mamba_b200 import generate, ops, synthetic, train
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
    torch.manual_seed(0)
    model = train.new_model("mamba").to(dev).eval()
    src, _, meta = synthetic.batch(10, 2048, seed=3)
    with torch.no_grad():
        dec = generate.RecurrentDecoder(model, 10, use_graph=True)
        dec.prefill(src.to(dev), meta.to(dev))
        for _ in range(8):
            dec.step()
        torch.cuda.synchronize()
        n = 200
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            dec.step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    out["decode"] = {"sequences": 10, "prompt": 2048, "timed_tokens_per_seq": n, "dtype": "f32", "ms_per_step": ms / n,
                     "tokens_per_sec": 10 * n / (ms * 1e-3), "mode": "recurrent step kernels, one CUDA graph per token, greedy"}
    return out


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
        "config": {"workload": WORKLOAD, "global_batch_tokens": B * T * world, "parallelism": f"dp{world}",
                   "cuda_graph": not args.no_graph,
                   "l2": "no explicit flush: one step streams ~1.4 GB of weights/grads/Adam state plus >2 GB of "
                         "activations, far above the 126 MB L2"},
        "clocks": clocks,
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps, "clocks": clocks_e2e},
        "gpu_launches": launches_per_step * args.steps,
        "gpu_launches_per_step": launches_per_step,
        "loss": loss_dev,
    }

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
        traffic = None
        try:
            if args.dtype == "bf16":
                traffic = json.loads((ROOT / "profiles" / "r01_ncu_traffic.json").read_text())[dom]["traffic_bytes"]
        except Exception:
            traffic = None
        line["roofline"] = {"kernel": dom, "bound": "hbm", "achieved": d["achieved_gbs"], "peak": peak, "unit": "GB/s",
                            "frac": d["frac_of_hbm_peak"], "traffic": traffic, "peak_source": peak_src,
                            "mean_us": d["mean_us"], "algorithmic_bytes": d["algorithmic_bytes"],
                            "binding_pipe": mufu_view(dom, B, T + cc.N_META, p.d_inner, p.d_state, d["mean_us"]),
                            "note": "d_state=64 makes this kernel MUFU/FMA-bound, not HBM-bound (SURVEY.md F8); "
                                    "see profiles/ for the ncu pipe utilisation"}
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
        print(json.dumps(line), flush=True)
    if world > 1:
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    return 0


if __name__ == "__main__":
    sys.exit(main())
