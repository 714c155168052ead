"""Times CBCA per mode at C3 on the bench image and on a piece-wise constant (worst case: 13-pixel arms) image:
   python scripts/cbca_modes.py [D] [--exact]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import synth_pair, flat_pair
pkg = importlib.import_module("mc-cnn-python_b200")
pf, ffi = pkg.process_functional, pkg._ffi
H = W = 1024; D = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 192
vol = torch.randn((H, W, D), device="cuda")
out = torch.empty_like(vol); scr = torch.empty_like(vol)
for image, maker in (("natural", synth_pair), ("flat", flat_pair)):
    li, ri = maker(H, W, 37, seed=0)
    arms, count = pf.cross_arms(li, 0.02, 14)
    a = arms.cpu().numpy().reshape(-1, 4)
    print("%s image: mean arms up/down/left/right %s, frac == 0 %.3f, frac > 2 %.3f, mean |U| %.1f, max |U| %d" % (
        image, a.mean(0).round(2), (a == 0).mean(), (a > 2).mean(), float(count.float().mean()), int(count.max())), flush=True)
    def bench(mode, name, rounds=16):
        def run(iters):
            ffi.call("mccnn_cbca", ffi.ptr(vol), ffi.ptr(out), ffi.ptr(scr), ffi.ptr(arms), ffi.ptr(count), D, H, W, iters, 14, mode, ffi.stream_ptr())
        run(4); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(rounds); b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / rounds
        print("  %-52s %.3f ms per round (call of %d)  %.0f GB/s (8 B/cell)" % (name, ms, rounds, 8.0 * H * W * D / ms / 1e6), flush=True)
    bench(0, "separable, chained rounds (default)")
    bench(0, "separable, chained rounds, call of 2", rounds=2)
    bench(2, "separable, two streaming passes per round")
    if "--exact" in sys.argv:
        bench(1, "exact (flat walk, bit-identical to reference)", rounds=2)
