"""Times the closing column pass with the winner-take-all folded in against the plain pass + k_wta (C3 volume)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import synth_pair
pkg = importlib.import_module("mc-cnn-python_b200")
pf, ffi = pkg.process_functional, pkg._ffi
H = W = 1024; D = 192
vol = torch.randn((H, W, D), device="cuda"); out = torch.empty_like(vol); scr = torch.empty_like(vol)
li, ri = synth_pair(H, W, 37, seed=0)
arms, count = pf.cross_arms(li, 0.02, 14)
keys = torch.empty((H, W), dtype=torch.int64, device="cuda"); disp = torch.empty((H, W), device="cuda")
def fused():
    ffi.call("mccnn_cbca_wta", ffi.ptr(vol), ffi.ptr(out), ffi.ptr(scr), ffi.ptr(arms), ffi.ptr(count), D, H, W, 2, 14, 0, 1, ffi.ptr(keys), ffi.ptr(disp), ffi.stream_ptr())
def split():
    ffi.call("mccnn_cbca", ffi.ptr(vol), ffi.ptr(out), ffi.ptr(scr), ffi.ptr(arms), ffi.ptr(count), D, H, W, 2, 14, 0, ffi.stream_ptr())
    ffi.call("mccnn_wta", ffi.ptr(out), ffi.ptr(disp), D, H, W, ffi.stream_ptr())
for name, fn in (("rows + colrow + close with WTA + decode", fused), ("rows + colrow + close + k_wta", split)) * 2:
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): fn()
    b.record(); torch.cuda.synchronize()
    print("%-32s %.3f ms" % (name, a.elapsed_time(b) / 20))
