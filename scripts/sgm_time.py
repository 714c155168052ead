"""Times the four chained SGM passes on both volumes at C3."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import synth_pair
pkg = importlib.import_module("mc-cnn-python_b200")
pf, ffi = pkg.process_functional, pkg._ffi
H = W = 1024; D = 192
li, ri = synth_pair(H, W, 37, seed=0)
il, ir = pf._image2d(li), pf._image2d(ri)
L = torch.randn((H, W, D), device="cuda"); R = torch.randn((H, W, D), device="cuda")
flags = pf._sgm_scratch(H, W, D)
def run():
    ffi.call("mccnn_sgm_average_pair", ffi.ptr(L), ffi.ptr(R), ffi.ptr(il), ffi.ptr(ir), ffi.ptr(flags), D, H, W, 2.3, 55.9, 4.0, 8.0, 0.08, 1.5, ffi.stream_ptr())
for _ in range(2): run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): run()
b.record(); torch.cuda.synchronize()
print("SGM x4 on two volumes: %.3f ms" % (a.elapsed_time(b) / 10))
