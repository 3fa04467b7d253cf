#!/usr/bin/env python
"""Summarise the gzipped `ncu --page raw --csv` exports that tools/ncu_capture.sh / ncu_multi.sh bring back from the GPU
box into the text files kept under profiles/ (one block per distinct kernel; launches of the same kernel and grid are
averaged).

    python tools/ncu_csv_summary.py gpurun_out/prof_x.raw.csv.gz [...] > profiles/r02_x.txt
"""
import collections
import csv
import gzip
import io
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_active.avg", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]
STALLS = "smsp__pcsamp_warps_issue_stalled_"


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return None


def main():
    for path in sys.argv[1:]:
        rows = list(csv.reader(io.TextIOWrapper(gzip.open(path))))
        hdr, units = rows[0], rows[1]
        groups = collections.OrderedDict()
        for vals in rows[2:]:
            rec = dict(zip(hdr, vals))
            groups.setdefault((rec.get("Kernel Name", "?"), rec.get("launch__grid_size")), []).append(rec)
        print(f"==== {path}")
        for (name, grid), recs in groups.items():
            print(f"== {name[:130]}   [{len(recs)} launch(es) averaged]")
            for k in KEYS:
                vs = [num(r[k]) for r in recs if k in r and num(r[k]) is not None]
                if vs:
                    print(f"  {k:72s} {sum(vs) / len(vs):16.3f} {units[hdr.index(k)]}")
            st = collections.Counter()
            for r in recs:
                for h, v in r.items():
                    if h.startswith(STALLS) and "not_issued" not in h and num(v):
                        st[h[len(STALLS):]] += num(v)
            tot = sum(st.values()) or 1.0
            print("  warp-state samples: " + ", ".join(f"{n} {100 * s / tot:.0f}%" for n, s in st.most_common(8)))
            rd = [num(r.get("dram__bytes_read.sum", "")) for r in recs]
            wr = [num(r.get("dram__bytes_write.sum", "")) for r in recs]
            if all(v is not None for v in rd + wr):
                u = units[hdr.index("dram__bytes_read.sum")]
                print(f"  dram traffic (read+write) per launch: {(sum(rd) + sum(wr)) / len(recs):.3f} {u}")


if __name__ == "__main__":
    main()
