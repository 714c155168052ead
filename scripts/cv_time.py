"""Times the cost-volume stage (tensor-core kernel + fill) at C3 and checks it against the previous result file if given."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import unit_features
pkg = importlib.import_module("mc-cnn-python_b200")
ffi = pkg._ffi
H, W, D = 1024, 1024, 192
fl, fr = (torch.from_numpy(x).cuda() for x in unit_features(H, W))
L = torch.empty(H, W, D, device="cuda"); R = torch.empty_like(L)
def run():
    ffi.call("mccnn_cost_volume", ffi.ptr(fl), ffi.ptr(fr), ffi.ptr(L), ffi.ptr(R), H, W, 64, D, ffi.stream_ptr())
for _ in range(3): run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): run()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
print("cost volume + fill: %.3f ms per call = %.0f GB/s algorithmic (%.1f %% of 6549)" % (ms, (8 + 512 / D) * H * W * D / ms / 1e6, (8 + 512 / D) * H * W * D / ms / 1e6 / 65.49))
