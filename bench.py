#!/usr/bin/env python
"""bench.py -- cost-volume cells per second of the stereo-matching hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c1|c4|c5|c5s] [--impl ours|reference]

Workloads c5 / c5s are ONE pair shared by all ranks (disparity-slab partition with NVLink re-partitioning around
SGM, mc-cnn-python_b200/slab.py; strong scaling); the others are one pair per rank.  With --gpus N > 1 and the default
workload the line also carries `slab` (the C5 pair shared by the N ranks, checked bit for bit against the single-GPU
map) and `c4` (one Middlebury-shaped pair per rank).

One "step" = one pass of the hot path (match.py:131-175: features -> cost volume -> CBCA x2 ->
4 chained SGM passes -> CBCA x16 -> WTA -> LR-check/interpolation -> sub-pixel -> median ->
bilateral) over one synthetic stereo pair per rank.  metric = H*W*D cells per second, whole job
(all ranks; image-pair data parallel, no collective on the data path => weak scaling).

  value  : inputs (two normalised images) already resident in HBM, CUDA-event timed.
  e2e    : the same step through StereoMatcher.run_host: NumPy images in, NumPy disparity out,
           pinned H2D and D2H copies inside the timed region.
  --impl reference : the reference's CPU path on the box's host cores.  The reference is Python 2 + TensorFlow; what can
           run here is (a) its C restatement (oracle/mccnn_oracle.c, pinned bit-exact to the reference's NumPy code)
           with every host thread -- the line's value: one full frame, then row samples -- and (b) the reference's OWN
           process_functional.py (py3-patched copy staged in oracle/_ref by oracle/stage_ref.py) on a small sample,
           reported beside it as cpu_baseline.reference_python.
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
def _oracle(threads=None):
    """The C restatement, built, with an EXPLICIT thread count (torchrun exports OMP_NUM_THREADS=1)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    O.build()
    O.set_threads(threads or os.cpu_count() or 1)
    return O


def bench_config(workload, image="natural"):
    """The `config` object of a line: the same for the GPU arm and the reference arm."""
    H, W, D, stages, desc = WORKLOADS[workload]
    vol_mb = H * W * ((D + 3) // 4 * 4) * 4 / 1e6
    return {"workload": desc, "H": H, "W": W, "D": D, "image": image, "weights": "random-init (glorot-uniform, seed 0)",
            "l2": "no explicit flush: every stage streams whole cost volumes (%.0f MB each) >> 126 MB L2" % vol_mb
                  if vol_mb > 300 else "volumes of %.0f MB: L2 flushed between timed steps by a 256 MB write" % vol_mb}


def cpu_pipeline_rate(H, W, D, stages, threads=None, seed=0):
    """Time the C restatement of the reference on (H, W, D); returns (cells/s, seconds, threads)."""
    O = _oracle(threads)
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


def reference_python_rate(D, stages, target_s=30.0):
    """The reference's OWN code (oracle/_ref/process_functional.py, staged by oracle/stage_ref.py from /root/reference)
    on a small sample with the workload's D: post-CNN stages a3-a12 exactly as match.py:137-175 calls them (the CNN is
    TensorFlow 1.x and cannot run; features come from the C port and are outside the timer).  Single threaded, like the
    reference.  Returns None when the staged copy is absent."""
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.isfile(os.path.join(ref_dir, "process_functional.py")):
        return None
    code = r"""
import sys, time, json, contextlib, io
import numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r); sys.path.insert(1, %r)
import process_functional as P                       # oracle/_ref: the reference's file, py3-patched
import oracle as O
from bench import synth_pair, unit_features
H, W, D, full = %d, %d, %d, %r
li, ri = synth_pair(H, W, min(16, max(1, D // 4)), 0)
if full:
    ws, bs = O.glorot_uniform_weights(seed=0)
    fl, fr = O.compute_features(li, ri, 11, 11, (ws, bs))
else:
    fl, fr = unit_features(H, W, 0)
t0 = time.perf_counter()
with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
    L, R = P.compute_cost_volume(fl, fr, D)
    if full:
        L, R = P.cost_volume_aggregation(li, ri, L, R, 0.02, 14, 2)
        L, R = P.SGM_average(L, R, li, ri, 2.3, 55.9, 4, 8, 0.08, 1.5)
        L, R = P.cost_volume_aggregation(li, ri, L, R, 0.02, 14, 16)
    dl, dr = P.disparity_prediction(L, R)
    if full:
        d = P.interpolation(dl, dr, D)
        d = P.subpixel_enhance(d, L)
        d = P.median_filter(d, 5, 5)
        d = P.bilateral_filter(li, d, 5, 5, 0, 6, 2)
dt = time.perf_counter() - t0
print(json.dumps({"cells": H * W * D, "seconds": dt}))
"""
    full = stages is None
    W = D + 2 + 14                       # narrowest legal width for this D, plus room for an arm
    # the full pipeline runs at ~1e4 cells/s (8.6e3 at 128x128x32), cost volume + WTA at ~1e6
    rate_guess = 1.0e4 if full else 1.0e6
    H = int(max(8, min(64, target_s * rate_guess / (W * D))))
    src = code % (ref_dir, os.path.join(ROOT, "oracle"), ROOT, H, W, D, full)
    try:
        r = subprocess.run([sys.executable, "-c", src], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=240,
                           env=dict(os.environ, OMP_NUM_THREADS="1"))
        out = json.loads(r.stdout.strip().splitlines()[-1])
    except Exception as e:                                   # the baseline beside the port, never a reason to fail the arm
        return {"unavailable": "reference run failed: %s" % (repr(e)[:200])}
    return {"value": out["cells"] / out["seconds"], "unit": UNIT, "cores": 1, "kind": "reference",
            "sample": "%dx%dx%d, %s, %.1f s, /root/reference/src/process_functional.py (py3-patched copy in oracle/_ref), "
                      "single threaded" % (H, W, D, "post-CNN stages a3-a12 (features outside the timer)" if full
                                           else "cost volume + WTA", out["seconds"])}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    H, W, D, stages, desc = WORKLOADS[args.workload]
    threads = os.cpu_count() or 1
    budget = float(os.environ.get("MCCNN_REF_BUDGET_S", "120"))            # whole arm, seconds of CPU pipeline time
    # step 1 is the FULL frame; the other steps are row samples (same W and D) sized to share what is left of the budget
    rate32, _, threads = cpu_pipeline_rate(min(H, 32), W, D, stages, threads)
    for _ in range(max(args.warmup - 1, 0)):
        cpu_pipeline_rate(min(H, 32), W, D, stages, threads)
    _, dt_full, threads = cpu_pipeline_rate(H, W, D, stages, threads)
    cells, dt = float(H) * W * D, dt_full
    rest = max(args.steps - 1, 0)
    hs = H
    if rest:
        per_step = max(budget - dt_full, 0.0) / rest
        hs = int(max(min(H, 32), min(H, (cells / dt_full) * per_step / (W * D))))
        for _ in range(rest):
            dt += cpu_pipeline_rate(hs, W, D, stages, threads)[1]
            cells += float(hs) * W * D
    value = cells / dt
    sample = ("step 1: the full %dx%dx%d frame in %.1f s (%.3g cells/s); steps 2..%d: %d-row samples of it (same W, D)"
              % (H, W, D, dt_full, H * W * D / dt_full, args.steps, hs))
    base = {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
            "full_frame": {"seconds": dt_full, "value": H * W * D / dt_full}}
    refpy = reference_python_rate(D, stages)
    if refpy is not None:
        base["reference_python"] = refpy
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * H * W * D / value,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(args.workload), "cpu_baseline": base,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "value: the C restatement of the reference (oracle/mccnn_oracle.c, OpenMP, %d threads set explicitly); "
                    "ms_per_step is per full frame at that rate; the reference's own Python is in "
                    "cpu_baseline.reference_python" % threads}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------ GPU arm
class Timer(object):
    """W warm-up steps, then K steps between a barrier + synchronize on both sides, CUDA events on the launching stream,
    max over ranks.  Small workloads flush L2 between timed steps (each step gets its own event pair)."""

    def __init__(self, torch, dist, world):
        self.torch, self.dist, self.world = torch, dist, world
        self.flush = None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        if self.world == 1:
            return list(vals)
        t = self.torch.tensor(list(vals), device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def run(self, fn, steps, warmup, flush_l2=False):
        torch = self.torch
        if flush_l2 and self.flush is None:
            self.flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda")        # 256 MB > 126 MB L2
        for _ in range(max(warmup, 3)):
            fn()
        self.barrier()
        t0 = time.time()
        if not flush_l2:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(steps):
                fn()
            b.record()
            self.barrier()
            ms = a.elapsed_time(b)
        else:
            evs = []
            for _ in range(steps):
                self.flush.fill_(0.0)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                evs.append((a, b))
            self.barrier()
            ms = sum(a.elapsed_time(b) for a, b in evs)
        return ms, t0, time.time()


def slab_leg(pkg, timer, rank, world, steps, warmup):
    """The C5 pair (2000x3000x400) shared by all ranks by disparity slab (strong scaling), and the proof that the
    partition changes nothing: the slab map against the single-GPU StereoMatcher map of the same pair on rank 0."""
    import torch
    H, W, D, _, desc = WORKLOADS["c5"]
    li, ri = synth_pair(H, W, min(37, D // 4), seed=0)
    cells = float(H) * W * D
    out = {"workload": desc, "H": H, "W": W, "D": D}
    m = pkg.SlabMatcher(H, W, D, checkpoint=None)
    m.set_images(li, ri)
    ms, _, _ = timer.run(m.run, steps, warmup)
    ms = timer.max_over_ranks(ms)[0]
    acc = {}
    for _ in range(2):
        for k, v in m.run_timed().items():
            acc[k] = acc.get(k, 0.0) + v / 2
    d_slab = m.run().clone()
    transport = m.transport
    del m
    torch.cuda.empty_cache()
    timer.barrier()
    parity = ms1 = None
    if rank == 0:                                            # the other ranks wait at the barrier below
        one = pkg.StereoMatcher(H, W, D, checkpoint=None)
        one.set_images(li, ri)
        for _ in range(2):
            one.run()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        d_one = one.run()
        b.record()
        torch.cuda.synchronize()
        ms1 = a.elapsed_time(b)
        parity = bool(torch.equal(d_one, d_slab))
        out["map_cells_differing"] = int((d_one != d_slab).sum().item())
        del one
        torch.cuda.empty_cache()
    timer.barrier()
    out.update({"ms_per_pair": ms / steps, "cells_per_s": cells * steps / (ms * 1e-3), "transport": transport,
                "phases_ms": acc, "single_gpu_ms_per_pair": ms1,
                "speedup_vs_n1": (ms1 / (ms / steps)) if ms1 else None, "parity": parity,
                "parity_check": "torch.equal(slab map, single-GPU StereoMatcher map of the same pair), rank 0"})
    return out


def c4_leg(pkg, timer, rank, world, steps, warmup):
    """BASELINE config 4: one Middlebury-v3 half-res shaped pair (1000x1500x256) per rank, no collective."""
    H, W, D, _, desc = WORKLOADS["c4"]
    m = pkg.StereoMatcher(H, W, D, checkpoint=None)
    li, ri = synth_pair(H, W, min(37, D // 4), seed=rank)
    m.set_images(li, ri)
    ms, _, _ = timer.run(m.run, steps, warmup)
    ms = timer.max_over_ranks(ms)[0]
    return {"workload": desc, "H": H, "W": W, "D": D, "pairs_per_step": world, "ms_per_step": ms / steps,
            "cells_per_s": world * float(H) * W * D * steps / (ms * 1e-3)}


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
    timer = Timer(torch, dist, world)
    H, W, D, stages, desc = WORKLOADS[args.workload]
    full = stages is None
    single_pair = args.workload in SINGLE_PAIR
    slab = single_pair and world > 1
    if slab:
        m = pkg.SlabMatcher(H, W, D, checkpoint=None)
    else:
        m = pkg.StereoMatcher(H, W, D, checkpoint=None, **({} if full else {"stages": stages}))
    maker = flat_pair if args.image == "flat" else synth_pair
    li, ri = maker(H, W, min(37, D // 4), seed=0 if single_pair else rank)      # one pair per rank / one for all
    m.set_images(li, ri)
    feats = None
    if not full:
        feats = unit_features(H, W, seed=rank)
        m.set_features(*feats)
    small = H * W * ((D + 3) // 4 * 4) * 4 < 300e6            # volumes that could sit in L2: flush between timed steps

    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = ffi.launch_count()
    ms, t0, t1 = timer.run(m.run, args.steps, args.warmup, flush_l2=small)
    launches = (ffi.launch_count() - launches0) * args.steps // (args.steps + max(args.warmup, 3))
    clocks = sampler.stop(t0, t1) if sampler else None

    # end to end through the host-buffer API (pinned H2D + D2H every step)
    if full:
        e2e_fn = lambda: m.run_host(li, ri)
        h2d, d2h = 2 * H * W * 4, H * W * 4
    else:
        # cost volume + WTA: the stage set's real inputs are the two feature maps, its outputs the two WTA maps
        pin = [torch.from_numpy(np.ascontiguousarray(f)).pin_memory() for f in feats]
        pout = [torch.empty((H, W), dtype=torch.float32).pin_memory() for _ in range(2)]

        def e2e_fn():
            for i in range(2):
                m.feat[i].copy_(pin[i], non_blocking=True)
            m.run()
            for i in range(2):
                pout[i].copy_(m.disp[i], non_blocking=True)
            torch.cuda.current_stream().synchronize()
        h2d, d2h = 2 * H * W * 64 * 4, 2 * H * W * 4
    e2e_fn()
    timer.barrier()
    te0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_fn()
    torch.cuda.synchronize()
    e2e_ms = 1e3 * (time.perf_counter() - te0)
    ms, e2e_ms = timer.max_over_ranks(ms, e2e_ms)

    cells = float(H) * W * D
    pairs = 1 if single_pair else world
    value = pairs * cells * args.steps / (ms * 1e-3)
    e2e_value = pairs * cells * args.steps / (e2e_ms * 1e-3)

    # per-stage device times (CUDA events on the launching stream), averaged over a few passes; the slab
    # partition's passes are collective, so every rank makes them (rank 0's times are reported)
    reps = 3
    acc = {}
    if rank == 0 or slab:
        for _ in range(reps):
            for k, v in m.run_timed().items():
                acc[k] = acc.get(k, 0.0) + v / reps
    # worst case for the aggregation: the same step on a piece-wise constant pair, every arm at its limit
    acc_flat = {}
    if rank == 0 and world == 1 and full and args.image != "flat" and not single_pair and not args.no_flat_stage:
        m.set_images(*flat_pair(H, W, min(37, D // 4), seed=0))
        m.run()
        for _ in range(reps):
            for k, v in m.run_timed().items():
                acc_flat[k] = acc_flat.get(k, 0.0) + v / reps
        m.set_images(li, ri)

    # with several GPUs the default run also measures the two multi-GPU configurations of BASELINE.json
    extra = {}
    if world > 1 and args.workload == "c3" and not args.no_extra_legs:
        del m
        torch.cuda.empty_cache()
        timer.barrier()
        extra["slab"] = slab_leg(pkg, timer, rank, world, max(3, min(args.steps, 5)), 3)
        extra["c4"] = c4_leg(pkg, timer, rank, world, max(3, min(args.steps, 5)), 3)
        m = None

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        hp = pkg.pipeline.DEFAULTS
        it1, it2 = int(hp["cbca_num_iterations1"]), int(hp["cbca_num_iterations2"])
        cells_rank = cells / world if slab else cells           # a rank's slab of every volume
        # algorithmic bytes per stage and launches of the stage's main kernel
        # ("launch" of a CBCA round = one round on one volume: a call of n rounds is  k_cbca_pass<rows> | (n-1) x
        #  k_cbca_colrow_g | k_cbca_pass<cols>, n + 1 kernels that each move 8 B/cell; the algorithmic figure is 8 B/cell/round)
        cbca_kernels = "k_cbca_colrow_g (+ k_cbca_pass<rows> / <cols> at the ends of a call; the closing <cols> of the second aggregation also takes the WTA minimum)"
        model = {
            "cbca2": (8.0 * cells_rank * it2 * 2, it2 * 2, cbca_kernels),
        } if slab else {
            "cost_volume": ((8.0 + 512.0 / D) * cells, 1, "k_cost_volume_tc (+k_cost_fill)"),
            "cbca1": (8.0 * cells * it1 * 2, it1 * 2, cbca_kernels),
            "sgm": (8.0 * cells * 4 * 2, 4, "k_sgm_pass"),
            "cbca2": (8.0 * cells * it2 * 2, it2 * 2, cbca_kernels),
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
            kernels["features"] = {"kernel": "k_conv64_h (tcgen05 kind::f16, fp16 hi/lo x3 products = float32 accuracy)", "ms": acc["features"], "TFLOPs": fl / 1e12,
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
            # what the kernels of one round actually move (ncu) against the same peak
            roofline["traffic_frac"] = roofline["traffic"] / (roofline["avg_launch_ms"] * 1e-3) / 1e9 / peak
        cfg = bench_config(args.workload, args.image)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong" if single_pair else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": cfg,
                "run": {"pairs_per_step": pairs,
                        "parallelism": ("one pair, %d disparity slabs; row / column slabs around SGM (3 NVLink "
                                        "re-partitions per volume, transport %s)" % (world, m.transport)) if slab else
                                       "image-pair data parallel, dp%d, no data-path collective" % world},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_ms / args.steps},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
                "stages_ms": acc, "kernels": kernels}
        if acc_flat:
            line["stages_ms_flat_image"] = acc_flat
            line["cbca_ms_per_round_per_volume"] = {"natural": acc["cbca2"] / (2 * it2), "flat": acc_flat["cbca2"] / (2 * it2),
                                                    "note": "flat = piece-wise constant pair, 13-pixel arms, regions up to 729 (pf:585-599)"}
        line.update(extra)
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
    ap.add_argument("--image", default="natural", choices=["natural", "flat"],
                    help="flat: piece-wise constant pair, every cross arm at its limit (worst case for the aggregation)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flat-stage", action="store_true", help="skip the extra passes on the piece-wise constant image")
    ap.add_argument("--no-extra-legs", action="store_true", help="with --gpus N > 1: skip the slab (c5) and c4 legs")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
