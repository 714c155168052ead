#!/usr/bin/env python
"""Per-source-line executed instructions / stall samples from an .ncu-rep captured with --import-source on.
    python scripts/ncu_lines.py gpurun_out/x.ncu-rep [min_pct]"""
import collections, csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
minp = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
rows = list(csv.reader(io.StringIO(out)))
sections, cur = [], None
for r in rows:
    if r and r[0] == "File Path":
        cur = {"file": r[1], "rows": []}
        sections.append(cur)
        continue
    if cur is not None:
        cur["rows"].append(r)
for sec in sections:
    rs = sec["rows"]
    hi = [i for i, r in enumerate(rs) if "Instructions Executed" in r]
    if not hi:
        continue
    hdr = rs[hi[0]]
    il, ie, iss, isrc = hdr.index("Line No"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
    agg = collections.OrderedDict()
    for r in rs[hi[0] + 1:]:
        if len(r) != len(hdr) or not r[il].strip():
            continue
        try:
            ln, n, s = int(r[il]), int(r[ie]), int(r[iss])
        except ValueError:
            continue
        a = agg.setdefault(ln, [0, 0, r[isrc]])
        a[0] += n
        a[1] += s
    tot = sum(v[0] for v in agg.values()) or 1
    ts = sum(v[1] for v in agg.values()) or 1
    print("== %s  warp-instructions %d  samples %d" % (sec["file"], tot, ts))
    for ln, v in sorted(agg.items()):
        if 100.0 * v[0] / tot >= minp or 100.0 * v[1] / ts >= minp:
            print("%4d inst %5.1f%% samp %5.1f%%  %s" % (ln, 100.0 * v[0] / tot, 100.0 * v[1] / ts, v[2].strip()[:110]))
