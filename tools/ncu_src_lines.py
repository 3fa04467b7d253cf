#!/usr/bin/env python
"""Join an `ncu --page source --csv` export (per-SASS-instruction warp-stall samples) with the line table of the same
kernel (nvdisasm -g of the cubin extracted from the built object) and print the samples per source line.

    python tools/ncu_src_lines.py gpurun_out/x.src.csv.gz <object-or-.so> <kernel-name-substring> [top]
"""
import collections, csv, gzip, io, re, subprocess, sys, tempfile, os

src, obj, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(io.TextIOWrapper(gzip.open(src))))
hdr = rows[1]
ia, isamp = hdr.index("Address"), hdr.index("# Samples")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
recs = rows[2:]
base = int(recs[0][ia], 16)
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")]
text = ""
for c in cubin:
    text += subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, c)], capture_output=True, text=True).stdout
cur, line, off2line = False, None, {}
for ln in text.splitlines():
    if ln.startswith(".text."):
        cur = kname in ln
        continue
    if not cur:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        line = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m:
        off2line[int(m.group(1), 16)] = line
by = collections.defaultdict(lambda: collections.Counter())
tot = 0
for r in recs:
    off = int(r[ia], 16) - base
    n = int(r[isamp] or 0)
    tot += n
    key = off2line.get(off, ("?", 0))
    by[key]["samples"] += n
    for i, h in stall_cols:
        v = int(r[i] or 0)
        if v:
            by[key][h] += v
print(f"total samples {tot}")
for key, c in sorted(by.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    st = ", ".join(f"{h[6:]} {v}" for h, v in c.most_common(5) if h != "samples")
    print(f"{key[0]}:{key[1]:<5d} {c['samples']:6d} ({100 * c['samples'] / tot:4.1f}%)  {st}")
