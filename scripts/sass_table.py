#!/usr/bin/env python
"""Per-kernel SASS opcode table of libmccnn_b200.so (what proves a Blackwell-native kernel: UTC*MMA = tcgen05.mma,
LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor load/store, UBLKCP = cp.async.bulk, LDGSTS = cp.async,
SYNCS = mbarrier, REDUX = redux.sync):   python scripts/sass_table.py > profiles/rNN_sass_opcodes.md"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "mc-cnn-python_b200", "libmccnn_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMAPF", "UTMASTG", "UBLKCP", "LDGSTS", "SYNCS", "REDUX", "CREDUX", "HMMA", "SHFL",
         "LDG", "STG", "LDS", "STS", "BAR"]
kern, counts, total = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern).replace("void ", "").replace("mccnn::", "")
        counts[kern] = collections.Counter(); total[kern] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and kern:
        op = m.group(1)
        total[kern] += 1
        for w in WATCH:
            if op == w or (w in ("UTCHMMA", "UTCQMMA") and op.startswith(w)):
                counts[kern][w] += 1
print("# SASS opcode counts per kernel (`cuobjdump -sass mc-cnn-python_b200/libmccnn_b200.so`, sm_100a)\n")
print("Static instruction counts.  UTCHMMA = `tcgen05.mma`, LDTM / STTM = `tcgen05.ld` / `tcgen05.st`, UTCBAR = `tcgen05.commit`,")
print("UTMALDG / UTMAPF / UTMASTG = TMA tensor load / L2 prefetch / store, LDGSTS = `cp.async`, SYNCS = mbarrier, REDUX / CREDUX = `redux.sync`.  No kernel uses the")
print("legacy `HMMA` path; no kernel uses a TMA store (`UTMASTG`): every result leaves through `STG`.\n")
cols = [w for w in WATCH if any(c[w] for c in counts.values()) or w in ("UTMASTG", "HMMA")]
print("| kernel | SASS instr. | " + " | ".join(cols) + " |")
print("|---|---|" + "---|" * len(cols))
for k, c in counts.items():
    print("| `%s` | %d | " % (k, total[k]) + " | ".join(str(c[w]) if c[w] else "" for w in cols) + " |")
