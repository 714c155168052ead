"""One GPU, all ranks emulated (LocalComm): where does the slab partition of a pair stop being bit-identical to the
single-GPU pipeline?   python scripts/slab_bisect.py H W D world [p2p]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from test_gpu_baseline_shapes import fast_pair
pkg = importlib.import_module("mc-cnn-python_b200")
H, W, D, world = (int(x) for x in sys.argv[1:5])
p2p = len(sys.argv) > 5 and sys.argv[5] == "p2p"
li, ri = fast_pair(H, W, 37, seed=1)
one = pkg.StereoMatcher(H, W, D); one.set_images(li, ri); want = one.run().clone()
plan = pkg.SlabPlan(H, W, D, world)
comm = pkg.LocalComm(world)
if p2p:
    arenas = comm.make_arenas(6 * plan.region_floats())
    ranks = [pkg.SlabRank(plan, r, arena=arenas[r]) for r in range(world)]
else:
    ranks = [pkg.SlabRank(plan, r) for r in range(world)]
for rk in ranks:
    rk.set_images(li, ri)
maps = (pkg.run_slabs_p2p if p2p else pkg.run_slabs)(ranks, comm)
torch.cuda.synchronize()
def cmp(name, a, b):
    same = bool(torch.equal(a, b))
    extra = ""
    if not same:
        ne = (a != b)
        idx = ne.nonzero()
        extra = " first %s last %s count %d maxdiff %.3g" % (idx[0].tolist(), idx[-1].tolist(), int(ne.sum()), float((a - b).abs().max()))
    print("%-28s %s%s" % (name, same, extra), flush=True)
    return same
for i in range(2):
    cmp("features[%d]" % i, ranks[0].feat[i], one.feat[i])
for r, rk in enumerate(ranks):
    b, c = plan.d_base(r), plan.d_count(r)
    for v in range(2):
        cmp("rank %d SGM result vol %d" % (r, v), rk.volB[v][:, :, :c], one.volB[v][:, :, b:b + c])
        cmp("rank %d CBCA2 result vol %d" % (r, v), rk.volA[v][:, :, :c], one.final_volume[v][:, :, b:b + c])
cmp("wta left", ranks[0].disp[0], one.disp[0]); cmp("wta right", ranks[0].disp[1], one.disp[1])
cmp("final map", maps[0], want)
