"""Times one CBCA round per mode at C3 (left volume of the bench pair)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import synth_pair
pkg = importlib.import_module("mc-cnn-python_b200")
pf, ffi = pkg.process_functional, pkg._ffi
H = W = 1024; D = 192
li, ri = synth_pair(H, W, 37, seed=0)
arms, count = pf.cross_arms(li, 0.02, 14)
ws = pf.cbca_workspace(H, W)
vol = torch.randn((H, W, D), device="cuda")
out = torch.empty_like(vol); scr = torch.empty_like(vol)
for mode, name in ((0, "separable (2 streaming passes)"), (2, "separable tiled (TMA)"), (1, "exact (flat walk, bit-identical)")):
    def run(iters=4):
        ffi.call("mccnn_cbca", ffi.ptr(vol), ffi.ptr(out), ffi.ptr(scr), ffi.ptr(arms), ffi.ptr(count), D, H, W, iters, 14, mode, ffi.ptr(ws), ffi.stream_ptr())
    run(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run(8); b.record(); torch.cuda.synchronize()
    print("%-40s %.3f ms per round" % (name, a.elapsed_time(b) / 8))
