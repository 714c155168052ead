"""Under torchrun on N GPUs: the NCCL disparity-slab partition of one pair against the single-GPU pipeline of the
same pair (each rank computes both; bit-identical maps expected), then phase times at a larger size.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/slab_check.py"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from bench import synth_pair
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pkg = importlib.import_module("mc-cnn-python_b200")
ok = True
for (H, W, D) in ((96, 200, 64), (130, 333, 100)):
    li, ri = synth_pair(H, W, 9, seed=1)
    one = pkg.StereoMatcher(H, W, D); one.set_images(li, ri); want = one.run().clone()
    for transport in ("nccl", "p2p"):
        sm = pkg.SlabMatcher(H, W, D, transport=transport); sm.set_images(li, ri)
        sm.run(); got = sm.run()
        torch.cuda.synchronize()
        same = bool(torch.equal(got, want))
        ok = ok and same
        print("rank %d: %dx%dx%d over %d ranks (%s): slab map == single-GPU map: %s" % (rank, H, W, D, world, transport, same), flush=True)
        del sm
    del one
H, W, D = 1024, 1536, 256
li, ri = synth_pair(H, W, 37, seed=0)
for transport in ("nccl", "p2p"):
    sm = pkg.SlabMatcher(H, W, D, transport=transport); sm.set_images(li, ri)
    for _ in range(2):
        sm.run()
    t = sm.run_timed()
    if rank == 0:
        print("phases (ms) at %dx%dx%d over %d ranks, %s:" % (H, W, D, world, transport), {k: round(v, 3) for k, v in t.items()}, "total %.2f" % sum(t.values()), flush=True)
    del sm
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(flag) == 1 else 1)
