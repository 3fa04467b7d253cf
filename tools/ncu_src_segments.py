#!/usr/bin/env python
"""Warp-stall samples of one kernel from an `ncu --page source --csv` export, summed per code segment between barriers
(BAR.SYNC / REDG = grid-barrier arrive / CCTL.IVALL = grid-barrier acquire): where a persistent, phase-structured
kernel spends its time.

    python tools/ncu_src_segments.py gpurun_out/x.src.csv.gz <kernel-name-substring> [top-instructions-per-segment]
"""
import collections, csv, gzip, io, sys

rows = list(csv.reader(io.TextIOWrapper(gzip.open(sys.argv[1]))))
want = sys.argv[2] if len(sys.argv) > 2 else ""
ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 5
# sections: a "Kernel Name" row, a header row, then one row per instruction
secs, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "recs": []}
        secs.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["recs"].append(r)
sec = next(s for s in secs if want in s["name"])
hdr, recs = sec["hdr"], sec["recs"]
print(sec["name"][:120])
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
segs = []
new = lambda k: {"start": k, "n": 0, "samp": 0, "ex": 0, "st": collections.Counter(), "top": []}
cur = new(0)
base = int(recs[0][ia], 16)
for k, r in enumerate(recs):
    s = r[isrc].strip()
    n, ex = int(r[isamp] or 0), int(r[iex] or 0)
    cur["n"] += 1; cur["samp"] += n; cur["ex"] += ex
    for i, h in stall:
        v = int(r[i] or 0)
        if v:
            cur["st"][h[6:]] += v
    cur["top"].append((n, hex(int(r[ia], 16) - base), s[:70]))
    if "REDG" in s or "CCTL.IVALL" in s or "BAR.SYNC" in s:
        cur["end"] = s[:40]; segs.append(cur); cur = new(k + 1)
cur["end"] = "END"; segs.append(cur)
tot = sum(s["samp"] for s in segs)
for s in segs:
    if s["samp"] < 0.004 * tot:
        continue
    print(f"--- instr {s['start']:5d}+{s['n']:<5d} executed {s['ex']:9d}  samples {s['samp']:6d} ({100 * s['samp'] / tot:4.1f}%)  ends: {s['end']}")
    print("     ", ", ".join(f"{k} {v}" for k, v in s["st"].most_common(6)))
    for n, a, t in sorted(s["top"], reverse=True)[:ntop]:
        print(f"        {n:5d} {a} {t}")
