"""Four CBCA rounds of one mode at C3 (for ncu): python scripts/cbca_one.py [mode] [D] [natural|flat]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import synth_pair, flat_pair
pkg = importlib.import_module("mc-cnn-python_b200")
pf, ffi = pkg.process_functional, pkg._ffi
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
H = W = 1024; D = int(sys.argv[2]) if len(sys.argv) > 2 else 192
maker = flat_pair if (len(sys.argv) > 3 and sys.argv[3] == "flat") else synth_pair
li, ri = maker(H, W, 37, seed=0)
arms, count = pf.cross_arms(li, 0.02, 14)
vol = torch.randn((H, W, D), device="cuda")
out = torch.empty_like(vol); scr = torch.empty_like(vol)
ffi.call("mccnn_cbca", ffi.ptr(vol), ffi.ptr(out), ffi.ptr(scr), ffi.ptr(arms), ffi.ptr(count), D, H, W, 4, 14, mode, ffi.stream_ptr())
torch.cuda.synchronize()
print("done")
