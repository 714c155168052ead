"""GPU: the chained CBCA rounds (MCCNN_CBCA_CHAIN=1) against the default two-pass rounds: identical bits on many shapes,
then ms per round at C3 on the natural and the flat image."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from bench import synth_pair, flat_pair
from test_gpu_parity import synth_images
pkg = importlib.import_module("mc-cnn-python_b200")
pf, ffi = pkg.process_functional, pkg._ffi
def cbca(vol, arms, count, D, H, W, iters, dist, chain):
    os.environ["MCCNN_CBCA_CHAIN"] = "1" if chain else "0"
    out = torch.empty_like(vol); scr = torch.empty_like(vol)
    ffi.call("mccnn_cbca", ffi.ptr(vol), ffi.ptr(out), ffi.ptr(scr), ffi.ptr(arms), ffi.ptr(count), D, H, W, iters, dist, 0, ffi.stream_ptr())
    torch.cuda.synchronize()
    return out
ok = True
cases = [(40, 90, 70, 4, 3, 14), (33, 47, 192, 30, 2, 14), (9, 29, 33, 1, 2, 14), (50, 21, 40, 2, 3, 14), (64, 64, 32, 1, 2, 14),
         (17, 200, 29, 3, 2, 14), (70, 40, 100, 2, 2, 14), (12, 129, 8, 1, 4, 14), (25, 300, 20, 1, 5, 14), (31, 191, 68, 2, 3, 20),
         (14, 140, 12, 1, 2, 40), (20, 65, 6, 1, 3, 3), (5, 64, 4, 2, 2, 1), (3, 128, 130, 1, 2, 14), (30, 62, 16, 1, 3, 14), (30, 63, 16, 1, 3, 14), (30, 125, 16, 2, 3, 14)]
for (H, W, D, levels, iters, dist) in cases:
    li, ri = synth_images(H * W + D, H, W, levels, 2)
    arms, count = pf.cross_arms(li, 0.02, dist)
    Dp = (D + 3) // 4 * 4
    vol = torch.randn((H, W, Dp), device="cuda") * 50
    a = cbca(vol, arms, count, D, H, W, iters, dist, False); b = cbca(vol, arms, count, D, H, W, iters, dist, True)
    same = bool(torch.equal(a[:, :, :D], b[:, :, :D]))
    ok = ok and same
    print((H, W, D, levels, iters, dist), "chained == two-pass:", same, flush=True)
H = W = 1024; D = 192
vol = torch.randn((H, W, D), device="cuda")
for image, maker in (("natural", synth_pair), ("flat", flat_pair)):
    li, ri = maker(H, W, 37, seed=0)
    arms, count = pf.cross_arms(li, 0.02, 14)
    a = cbca(vol, arms, count, D, H, W, 4, 14, False); b = cbca(vol, arms, count, D, H, W, 4, 14, True)
    same = bool(torch.equal(a, b)); ok = ok and same
    for chain in (False, True):
        os.environ["MCCNN_CBCA_CHAIN"] = "1" if chain else "0"
        out = torch.empty_like(vol); scr = torch.empty_like(vol)
        run = lambda it: ffi.call("mccnn_cbca", ffi.ptr(vol), ffi.ptr(out), ffi.ptr(scr), ffi.ptr(arms), ffi.ptr(count), D, H, W, it, 14, 0, ffi.stream_ptr())
        run(4); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(16); e1.record(); torch.cuda.synchronize()
        print("%s image, %s: %.3f ms per round (call of 16), same bits: %s" % (image, "chained" if chain else "two-pass", e0.elapsed_time(e1) / 16, same), flush=True)
print("ALL IDENTICAL" if ok else "MISMATCH")
