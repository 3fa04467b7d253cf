#!/usr/bin/env python
"""Summarise .ncu-rep captures (read here, no GPU needed) into the text files kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep [...] > profiles/r01_x.txt
"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_active.avg", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]
STALLS = "smsp__pcsamp_warps_issue_stalled_"


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            rec = dict(zip(hdr, vals))
            print(f"== {path} :: {rec.get('Kernel Name', '?')[:110]}")
            for k in KEYS:
                if k in rec:
                    print(f"  {k:75s} {rec[k]:>16s} {units[hdr.index(k)]}")
            stalls = sorted(((float(v), h[len(STALLS):]) for h, v in rec.items()
                             if h.startswith(STALLS) and "not_issued" not in h and v), reverse=True)
            tot = sum(s for s, _ in stalls) or 1.0
            print("  warp-state samples: " + ", ".join(f"{n} {100 * s / tot:.0f}%" for s, n in stalls[:8]))
            rd = float(rec.get("dram__bytes_read.sum", 0) or 0)
            wr = float(rec.get("dram__bytes_write.sum", 0) or 0)
            u = units[hdr.index("dram__bytes_read.sum")] if "dram__bytes_read.sum" in hdr else ""
            print(f"  dram traffic (read+write): {rd + wr:.3f} {u}")


if __name__ == "__main__":
    main()
