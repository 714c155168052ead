#!/usr/bin/env python
"""Turns ncu output brought back in gpurun_out/ into the small text summaries kept under profiles/.

    python scripts/ncu_summary.py launches gpurun_out/launches.csv  > profiles/rNN_launches.md
    python scripts/ncu_summary.py rep gpurun_out/prof_k.ncu-rep [...] > profiles/rNN_kernels.md
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__t_bytes.sum", "L1 bytes"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu pipe %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__shared_mem_per_block_static", "static smem/block"),
    ("launch__occupancy_limit_registers", "occ limit regs"),
    ("launch__occupancy_limit_shared_mem", "occ limit smem"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    start = next(i for i, r in enumerate(rows) if r[0] == "ID")
    hdr = rows[start]
    agg = collections.OrderedDict()
    for r in rows[start + 1:]:
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = d["Kernel Name"].split("(")[0].replace("void ", "")
        v = float(d["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(d["Metric Unit"], 1.0)
        a = agg.setdefault(k, [0, 0.0, d["Grid Size"], d["Block Size"]])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print("| kernel | launches | total us | avg us | share | grid | block |")
    print("|---|---|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("| %s | %d | %.1f | %.1f | %.3f | %s | %s |" % (k, v[0], v[1], v[1] / v[0], v[1] / tot, v[2], v[3]))
    print("\ntotal %.1f us over %d launches (ncu: cold-cache, serialised; compare shares, not absolutes)"
          % (tot, sum(v[0] for v in agg.values())))


def rep(paths):
    for p in paths:
        out = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], stdout=subprocess.PIPE,
                             stderr=subprocess.DEVNULL, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            print("## %s: unreadable" % p)
            continue
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            print("## %s  (%s)\n" % (d.get("Kernel Name", "?").split("(")[0], p.split("/")[-1]))
            for key, label in KEYS:
                if key in d:
                    print("- %s: %s %s   [`%s`]" % (label, d[key], u.get(key, ""), key))
            print()


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        rep(sys.argv[2:])
