#!/usr/bin/env python
"""bench.py -- cost-volume cells per second of the stereo-matching hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c1] [--impl ours|reference]

Workloads c5 / c5s are ONE pair shared by all ranks (disparity-slab partition with NVLink re-partitioning around
SGM, mc-cnn-python_b200/slab.py; strong scaling); the others are one pair per rank.

One "step" = one pass of the hot path (match.py:131-175: features -> cost volume -> CBCA x2 ->
4 chained SGM passes -> CBCA x16 -> WTA -> LR-check/interpolation -> sub-pixel -> median ->
bilateral) over one synthetic stereo pair per rank.  metric = H*W*D cells per second, whole job
(all ranks; image-pair data parallel, no collective on the data path => weak scaling).

  value  : inputs (two normalised images) already resident in HBM, CUDA-event timed.
  e2e    : the same step through StereoMatcher.run_host: NumPy images in, NumPy disparity out,
           pinned H2D and D2H copies inside the timed region.
  --impl reference : the reference's CPU path.  The reference is Python 2 + TensorFlow and cannot
           run (or travel) to the GPU box, so this arm times the C restatement of it
           (oracle/mccnn_oracle.c, pinned bit-exact to the reference's NumPy code) with all host
           threads, on a bounded sample of the same workload.
"""
import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (H, W, D, stages or None for the full pipeline, description)
    "c3": (1024, 1024, 192, None, "1024x1024x192 synthetic pair, full pipeline (features+CV+CBCA2+SGM4+CBCA16+WTA+refine)"),
    "c2": (512, 512, 128, ("cost_volume", "wta"), "512x512x128 synthetic pair, cost volume + WTA only"),
    "c1": (128, 128, 32, None, "128x128x32 synthetic pair, full pipeline"),
    "c4": (1000, 1500, 256, None, "Middlebury-v3 half-res shaped pairs (1000x1500x256), full pipeline, one pair per GPU"),
    # one big pair: a single GPU runs it whole, N > 1 GPUs share it by disparity slab (strong scaling)
    "c5": (2000, 3000, 400, None, "2000x3000x400 single pair, full pipeline, disparity-slab partition across the GPUs"),
    "c5s": (1024, 1536, 256, None, "1024x1536x256 single pair, full pipeline, disparity-slab partition across the GPUs"),
}
SINGLE_PAIR = ("c5", "c5s")
METRIC = "cost_volume_cells_per_sec"
UNIT = "cells/s"


# ------------------------------------------------------------------------------------------ inputs
def synth_pair(H, W, shift, seed=0):
    """Seeded natural-like stereo pair: multi-scale blurred noise, 2%/98% clipped (saturated flats),
    quantised to 8 bits, normalised exactly as match.py:118-123; right(x) = left(x + shift)."""
    rng = np.random.default_rng(seed)
    Wb = W + shift

    def blur(a, sigma):
        r = int(3 * sigma)
        k = np.exp(-0.5 * (np.arange(-r, r + 1) / sigma) ** 2)
        k /= k.sum()
        ap = np.pad(a, ((r, r), (r, r)), mode="reflect")
        ap = np.apply_along_axis(lambda v: np.convolve(v, k, mode="valid"), 0, ap)
        ap = np.apply_along_axis(lambda v: np.convolve(v, k, mode="valid"), 1, ap)
        return ap

    base = np.zeros((H, Wb))
    for sigma, amp in ((1.5, 1.0), (6.0, 2.0), (24.0, 4.0)):
        n = blur(rng.standard_normal((H, Wb)), sigma)
        base += amp * n / n.std()
    lo, hi = np.percentile(base, [2, 98])
    base = np.clip(base, lo, hi)
    q = np.floor((base - lo) / (hi - lo) * 255.0).astype(np.float32)
    left, right = q[:, :W], q[:, shift:shift + W]
    li = ((left - np.mean(left, axis=(0, 1))) / np.std(left, axis=(0, 1))).astype(np.float32)
    ri = ((right - np.mean(right, axis=(0, 1))) / np.std(right, axis=(0, 1))).astype(np.float32)
    return li[..., None], ri[..., None]


def flat_pair(H, W, shift, seed=0, block=48):
    """Worst case for cross-based aggregation: a piece-wise constant image of `block`-pixel squares of
    random 8-bit levels, so that almost every arm runs to the distance limit (13 pixels at match.py's
    distance_threshold 14, regions up to 27 x 27 = 729, pf:585-599); normalised as match.py:118-123,
    right(x) = left(x + shift)."""
    rng = np.random.default_rng(seed)
    Wb = W + shift
    lv = rng.integers(0, 256, ((H + block - 1) // block, (Wb + block - 1) // block)).astype(np.float32)
    q = np.kron(lv, np.ones((block, block), np.float32))[:H, :Wb]
    left, right = q[:, :W], q[:, shift:shift + W]
    li = ((left - np.mean(left, axis=(0, 1))) / np.std(left, axis=(0, 1))).astype(np.float32)
    ri = ((right - np.mean(right, axis=(0, 1))) / np.std(right, axis=(0, 1))).astype(np.float32)
    return li[..., None], ri[..., None]


def unit_features(H, W, seed=0):
    rng = np.random.default_rng(seed)
    f = rng.standard_normal((2, H, W, 64)).astype(np.float32)
    f /= np.linalg.norm(f, axis=-1, keepdims=True)
    return f[0], f[1]


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler(object):
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.thread = [], None, None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for _, r in self.rows]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except Exception:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_pipeline_rate(H, W, D, stages, threads=None, seed=0):
    """Time the C restatement of the reference on (H, W, D); returns (cells/s, seconds, threads)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    O.build()
    if threads:
        O.set_threads(threads)
    li, ri = synth_pair(H, W, min(16, max(1, D // 4)), seed)
    full = stages is None
    ws, bs = O.glorot_uniform_weights(seed=0)
    fl, fr = (None, None) if full else unit_features(H, W, seed)
    t0 = time.perf_counter()
    if full:
        fl, fr = O.compute_features(li, ri, 11, 11, (ws, bs))
        O.match_from_features(li, ri, fl, fr, D)
    else:
        L, R = O.compute_cost_volume(fl, fr, D)
        O.disparity_prediction(L, R)
    dt = time.perf_counter() - t0
    return H * W * D / dt, dt, O.num_threads()


def bounded_cpu_sample(H, W, D, stages, target_s=12.0):
    """Pick a sample (rows x W x D, same D and W) of about target_s seconds of CPU work."""
    h0 = min(H, 32)
    rate, dt, threads = cpu_pipeline_rate(h0, W, D, stages)
    h = int(max(h0, min(H, rate * target_s / (W * D))))
    return h, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    H, W, D, stages, desc = WORKLOADS[args.workload]
    hs, threads = bounded_cpu_sample(H, W, D, stages, target_s=6.0)
    for _ in range(args.warmup):
        cpu_pipeline_rate(min(hs, 32), W, D, stages)
    dt = 0.0                      # sum of the pipeline timers (input generation is outside them)
    for _ in range(args.steps):
        dt += cpu_pipeline_rate(hs, W, D, stages)[1]
    value = hs * W * D * args.steps / dt
    sample = "%dx%dx%d rows-sample of the %s workload per step (same W and D)" % (hs, W, D, args.workload)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "H": H, "W": W, "D": D},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference is Python2+TF1 (cannot run here); timed: its C restatement oracle/mccnn_oracle.c, OpenMP"}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pkg = importlib.import_module("mc-cnn-python_b200")
    ffi = pkg._ffi
    H, W, D, stages, desc = WORKLOADS[args.workload]
    full = stages is None
    single_pair = args.workload in SINGLE_PAIR
    slab = single_pair and world > 1
    if slab:
        m = pkg.SlabMatcher(H, W, D, checkpoint=None)
    else:
        m = pkg.StereoMatcher(H, W, D, checkpoint=None, **({} if full else {"stages": stages}))
    li, ri = synth_pair(H, W, min(37, D // 4), seed=0 if single_pair else rank)      # one pair per rank / one for all
    m.set_images(li, ri)
    if not full:
        m.set_features(*unit_features(H, W, seed=rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        m.run()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = ffi.launch_count()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.time()
    start.record()
    for _ in range(args.steps):
        m.run()
    end.record()
    barrier()
    t1 = time.time()
    launches = ffi.launch_count() - launches0
    ms = start.elapsed_time(end)
    clocks = sampler.stop(t0, t1) if sampler else None

    # end to end through the host-buffer API (pinned H2D + D2H every step)
    m.run_host(li, ri)
    barrier()
    te0 = time.perf_counter()
    for _ in range(args.steps):
        m.run_host(li, ri)
    torch.cuda.synchronize()
    e2e_ms = 1e3 * (time.perf_counter() - te0)

    if world > 1:
        t = torch.tensor([ms, e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(t[0]), float(t[1])

    cells = float(H) * W * D
    pairs = 1 if single_pair else world
    value = pairs * cells * args.steps / (ms * 1e-3)
    e2e_value = pairs * cells * args.steps / (e2e_ms * 1e-3)

    line = None
    # per-stage device times (CUDA events on the launching stream), averaged over a few passes; the slab
    # partition's passes are collective, so every rank makes them (rank 0's times are reported)
    reps = 3
    acc = {}
    if rank == 0 or slab:
        for _ in range(reps):
            for k, v in m.run_timed().items():
                acc[k] = acc.get(k, 0.0) + v / reps
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        hp = m.hp
        it1, it2 = int(hp["cbca_num_iterations1"]), int(hp["cbca_num_iterations2"])
        if slab:
            cells_rank = cells / world                       # this rank's slab of every volume
        # algorithmic bytes per stage and launches of the stage's main kernel
        # ("launch" of a CBCA round = its two streaming passes (k_cbca_pass, rows then columns) on one volume; the
        #  algorithmic figure is the fused minimum of 8 B/cell/round, the two passes actually move 16 B/cell)
        model = {
            "cbca2": (8.0 * cells_rank * it2 * 2, it2 * 2, "k_cbca_pass<rows>+k_cbca_pass<cols>"),
        } if slab else {
            "cost_volume": ((8.0 + 512.0 / D) * cells, 1, "k_cost_volume_tc (+k_cost_fill)"),
            "cbca1": (8.0 * cells * it1 * 2, it1 * 2, "k_cbca_pass<rows>+k_cbca_pass<cols>"),
            "sgm": (8.0 * cells * 4 * 2, 4, "k_sgm_pass"),
            "cbca2": (8.0 * cells * it2 * 2, it2 * 2, "k_cbca_pass<rows>+k_cbca_pass<cols>"),
            "wta": (4.0 * cells * 2, 2, "k_wta"),
        }
        kernels = {}
        for st, (nbytes, nl, kname) in model.items():
            if st in acc and acc[st] > 0:
                gbs = nbytes / (acc[st] * 1e-3) / 1e9
                kernels[st] = {"kernel": kname, "ms": acc[st], "launches": nl, "algorithmic_GB": nbytes / 1e9,
                               "achieved_GBps": gbs, "frac_of_hbm_peak": gbs / peak}
        if "features" in acc:
            fl = 2.0 * H * W * 296064.0
            kernels["features"] = {"kernel": "k_conv64_tc (tcgen05 tf32 x3)", "ms": acc["features"], "TFLOPs": fl / 1e12,
                                   "achieved_TFLOPps": fl / (acc["features"] * 1e-3) / 1e12}
        # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture
        traffic = {}
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        except Exception:
            pass
        dom = max((k for k in kernels if k in model), key=lambda k: acc[k])
        nbytes, nl, kname = model[dom]
        roofline = {"bound": "hbm", "kernel": kname, "stage": dom, "achieved": kernels[dom]["achieved_GBps"],
                    "peak": peak, "unit": "GB/s", "frac": kernels[dom]["achieved_GBps"] / peak,
                    "traffic": (traffic.get(dom) or {}).get("bytes_per_launch") if (H, W, D) == (1024, 1024, 192) else None,
                    "traffic_source": (traffic.get(dom) or {}).get("source"), "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": nbytes / nl, "avg_launch_ms": acc[dom] / nl}
        if roofline["traffic"]:
            # what the kernel(s) of one launch actually move (ncu) against the same peak: the default CBCA round is
            # two passes, i.e. twice the algorithmic bytes, and runs close to copy speed on those
            roofline["traffic_frac"] = roofline["traffic"] / (roofline["avg_launch_ms"] * 1e-3) / 1e9 / peak
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong" if single_pair else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": desc, "H": H, "W": W, "D": D, "pairs_per_step": pairs,
                           "parallelism": ("one pair, %d disparity slabs; row / column slabs around SGM (3 NVLink "
                                           "re-partitions per volume, transport %s)" % (world, m.transport)) if slab else
                                          "image-pair data parallel, dp%d" % world,
                           "weights": "random-init (glorot-uniform, seed 0)",
                           "l2": "no explicit flush: each stage streams >= 1.6 GB (volumes are 805 MB each) >> 126 MB L2"},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * H * W * 4,
                        "d2h_bytes_per_step": H * W * 4, "ms_per_step": e2e_ms / args.steps},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
                "stages_ms": acc, "kernels": kernels}
        if world == 1 and not args.no_cpu_baseline:
            hs, threads = bounded_cpu_sample(H, W, D, stages, target_s=12.0)
            rate, dt, threads = cpu_pipeline_rate(hs, W, D, stages)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "%dx%dx%d rows-sample of the workload (same W, D), %.1f s, "
                                              "C restatement of the reference (oracle/), OpenMP" % (hs, W, D, dt)}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
