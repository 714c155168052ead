"""Accuracy of the tensor-core cost volume against a float64 dot product (scale-relative), and timing."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import unit_features
pkg = importlib.import_module("mc-cnn-python_b200")
pf, ffi = pkg.process_functional, pkg._ffi
for (H, W, D) in [(64, 512, 128), (32, 1024, 192)]:
    fl, fr = unit_features(H, W, seed=1)
    L, R = pf.compute_cost_volume(fl, fr, D)
    fl64, fr64 = fl.astype(np.float64), fr.astype(np.float64)
    err = 0.0; scale = 0.0
    for d in range(0, D, 7):
        ref = -(fl64[:, d:, :] * fr64[:, :W - d, :]).sum(-1)
        err = max(err, np.abs(L[d][:, d:] - ref).max()); scale = max(scale, np.abs(ref).max())
        err = max(err, np.abs(R[d][:, :W - d] - ref).max())
    print("HxWxD %dx%dx%d: max abs err vs float64 %.3e, scale %.3f, relative %.3e" % (H, W, D, err, scale, err / scale))
H, W, D = 1024, 1024, 192
fl, fr = (torch.from_numpy(a).cuda() for a in unit_features(H, W))
Lh = torch.empty((H, W, D), device="cuda"); Rh = torch.empty_like(Lh)
def run():
    ffi.call("mccnn_cost_volume", ffi.ptr(fl), ffi.ptr(fr), ffi.ptr(Lh), ffi.ptr(Rh), H, W, 64, D, ffi.stream_ptr())
for _ in range(3): run()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): run()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
print("C3 cost volume (+fill): %.3f ms, %.0f GB/s algorithmic" % (ms, (8 + 512.0 / D) * H * W * D / ms / 1e6))
