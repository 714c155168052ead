"""Debug: dump the CBCA tile schedule built on the GPU and compare the warp-iteration count it implies with the ideal."""
import importlib, os, sys
import numpy as np
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
vol = torch.zeros((H, W, D), device="cuda")
out = torch.empty_like(vol)
ffi.call("mccnn_cbca", ffi.ptr(vol), ffi.ptr(out), None, ffi.ptr(arms), ffi.ptr(count), D, H, W, 1, 14, 0, ffi.ptr(ws), ffi.stream_ptr())
torch.cuda.synchronize()
R = 44
dt = np.dtype([("up", "u1"), ("down", "u1"), ("nrows", "u1"), ("gp", "u1"), ("r0", "<u2"), ("w0", "<u2"), ("h0", "<u2"),
               ("tw", "u1"), ("th", "u1"), ("stage_bytes", "<u4"), ("hwi", "u1", (R,)), ("rowb", "<u2", (R,)),
               ("centre", "<u2", (R,)), ("permA", "u1", (8, R)), ("permB", "u1", (8, 16)), ("arms", "u1", (42, 8, 4)), ("pad", "u1", (4,)),
               ("cnt", "<f4", (16, 8, 2))])
assert dt.itemsize == 3088, dt.itemsize
ntiles = (H // 16) * (W // 8)
raw = ws.cpu().numpy().view(np.uint8)[: ntiles * 3088]
m = raw.view(dt)
print("gp histogram", np.bincount(m["gp"]))
print("nrows mean", m["nrows"].mean(), "stage bytes mean", m["stage_bytes"].mean())
totA = totB = 0
for t in m[::7]:
    st = 5 if t["gp"] == 4 else (3 if t["gp"] == 2 else 1)
    nrows = int(t["nrows"])
    for step in range((nrows + 3) // 4):
        mx = 0
        for q in range(8):
            for rk in range(4):
                r = t["permA"][q][step * 4 + rk]
                if r == 255: continue
                px = (st * (q - int(t["centre"][r]))) & 7
                a = t["arms"][r][px]
                mx = max(mx, int(a[2]) + int(a[3]) + 1)
        totA += mx
    for step in range(4):
        mx = 0
        for q in range(8):
            for rk in range(4):
                rt = t["permB"][q][step * 4 + rk]
                if rt == 255: continue
                a = t["arms"][rt + int(t["up"])][q]
                mx = max(mx, int(a[0]) + int(a[1]) + 1)
        totB += mx
n = len(m[::7])
print("warp-iterations per tile from the GPU schedule: A %.1f B %.1f" % (totA / n, totB / n))
