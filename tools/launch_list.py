#!/usr/bin/env python
"""Turn an `ncu --metrics gpu__time_duration.sum --csv` log of bench.py into the per-kernel launch list of ONE
training step (kept under profiles/).  A step is the window between two launches of the loss-mean kernel (one per step).

    python tools/launch_list.py gpurun_out/launches.csv > profiles/rNN_bench_step_launches.txt
"""
import collections
import csv
import re
import sys


def short(name):
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)              # drop the parameter list
    name = re.sub(r"<.*$", "", name) if name.startswith(("at::", "native::", "at_cuda")) else name
    return name[:100]


def main():
    rows = []
    with open(sys.argv[1]) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((r["Kernel Name"], float(r["Metric Value"]) / 1e3))
    # one launch of the loss-mean kernel per step (the staged step interleaves its per-stage Adam launches with the
    # backward, so the optimizer is no delimiter): the window runs from one step's loss to the next one's
    marks = [i for i, (n, _) in enumerate(rows) if "loss_mean_kernel" in n]
    if len(marks) < 2:
        raise SystemExit("fewer than two training steps in the log")
    step = rows[marks[-2]:marks[-1]]
    agg = collections.OrderedDict()
    for n, us in step:
        k = short(n)
        c, t = agg.get(k, (0, 0.0))
        agg[k] = (c + 1, t + us)
    tot = sum(t for _, t in agg.values())
    print("ncu launch list of one training step (" + " ".join(sys.argv[2:]) + "; cold-cache, serialised: compare SHARES)")
    print(f"launches in step: {len(step)}; sum of durations: {tot:.1f} us\n")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t:9.1f} us {c:4d} {100 * t / tot:5.1f}%  {k}")


if __name__ == "__main__":
    main()
