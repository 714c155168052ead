"""Times one CBCA round per mode (and per strip shape of the marching kernel) at C3 (left volume of the bench pair)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import synth_pair
pkg = importlib.import_module("mc-cnn-python_b200")
pf, ffi = pkg.process_functional, pkg._ffi
H = W = 1024; D = int(sys.argv[1]) if len(sys.argv) > 1 else 192
li, ri = synth_pair(H, W, 37, seed=0)
arms, count = pf.cross_arms(li, 0.02, 14)
a = arms.cpu().numpy().reshape(H, W, 4)
print("arms: mean up/down/left/right", a.reshape(-1, 4).mean(0), "max", a.max(), "frac == 0", (a == 0).mean(), "frac > 2", (a > 2).mean())
ws = pf.cbca_workspace(H, W)
vol = torch.randn((H, W, D), device="cuda")
out = torch.empty_like(vol); scr = torch.empty_like(vol)
ref = None
def bench(mode, name):
    global ref
    def run(iters=4):
        ffi.call("mccnn_cbca", ffi.ptr(vol), ffi.ptr(out), ffi.ptr(scr), ffi.ptr(arms), ffi.ptr(count), D, H, W, iters, 14, mode, ffi.ptr(ws), ffi.stream_ptr())
    run(); torch.cuda.synchronize()
    if ref is None:
        ref = out.clone()
    same = bool(torch.equal(ref, out))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run(8); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 8
    print("%-44s %.3f ms per round  %.0f GB/s (8 B/cell)  same=%s" % (name, ms, 8.0 * H * W * D / ms / 1e6, same), flush=True)
bench(0, "separable (2 streaming passes)")
bench(4, "separable, row sums in L2 (persistent)")
for v in range(6):
    os.environ["MCCNN_CBCA_MARCH"] = "%d,0" % v
    bench(3, "march variant %d" % v)
os.environ.pop("MCCNN_CBCA_MARCH", None)
if "--all" in sys.argv:
    bench(2, "separable tiled (TMA)")
    bench(1, "exact (flat walk, bit-identical)")
