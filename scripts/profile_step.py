"""Runs the C3 hot path a few times (for ncu): python scripts/profile_step.py [steps] [workload]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import synth_pair, WORKLOADS, unit_features
pkg = importlib.import_module("mc-cnn-python_b200")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
wl = sys.argv[2] if len(sys.argv) > 2 else "c3"
H, W, D, stages, _ = WORKLOADS[wl]
m = pkg.StereoMatcher(H, W, D, **({} if stages is None else {"stages": stages}))
li, ri = synth_pair(H, W, min(37, D // 4), seed=0)
m.set_images(li, ri)
if stages is not None:
    m.set_features(*unit_features(H, W))
for _ in range(steps):
    m.run()
torch.cuda.synchronize()
print("done", pkg._ffi.launch_count())
