#!/usr/bin/env python
"""Top SASS instructions by stall samples, with the dominant stall reasons, from an .ncu-rep.
    python scripts/ncu_sass_stalls.py gpurun_out/x.ncu-rep [top_n]"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(io.StringIO(out)))
hdr = next(r for r in rows if r and r[0] == "Address")
isrc, isamp, iexec = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for n, r in enumerate(rows):
    if len(r) != len(hdr) or r[0] == "Address":
        continue
    try:
        s = int(r[isamp])
    except ValueError:
        continue
    reasons = sorted(((int(r[i] or 0), h[6:]) for i, h in stall), reverse=True)[:3]
    data.append((s, n, r[isrc].strip(), int(r[iexec] or 0), reasons))
total = sum(d[0] for d in data)
print("total samples", total)
for s, n, src, ex, reasons in sorted(data, reverse=True)[:top]:
    print("%5.1f%%  #%-5d exec %-9d %-60s %s" % (100.0 * s / total, n, ex, src[:60], " ".join("%s:%d" % (h, c) for c, h in reasons if c)))
