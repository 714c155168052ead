"""CPU, world_size 2 and 3 over gloo: the plumbing of the disparity-slab partition -- SlabPlan's
index arithmetic and DistComm's grouped send/recv, all-gather and all-reduce -- driven through the three
re-partitions d-slabs -> row slabs -> column slabs -> d-slabs with torch slicing standing in for the packing
kernel (the kernels themselves are GPU tests: tests/test_gpu_slab.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _volume(H, W, D):
    """cell (h, w, d) holds a unique number."""
    return torch.arange(H * W * D, dtype=torch.float32).view(H, W, D)


def _worker(rank, world, port, H, W, D, out_dir):
    sys.path.insert(0, ROOT)
    import importlib
    pkg = importlib.import_module("mc-cnn-python_b200")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    plan = pkg.SlabPlan(H, W, D, world)
    comm = pkg.DistComm()
    Dp = 4 * plan.G
    full = torch.zeros(H, W, Dp)
    full[:, :, :D] = _volume(H, W, D)
    g0, g1 = plan.granules[rank]
    slab = full[:, :, 4 * g0:4 * g1].contiguous()                               # what this rank owns at the start
    # d-slabs -> row slabs
    h0, h1 = plan.rows[rank]
    send = [slab[lo:hi].contiguous() for lo, hi in plan.rows]
    recv = [torch.empty(h1 - h0, W, 4 * plan.g_count(j)) for j in range(world)]
    comm.exchange([send], [recv])
    rows = torch.cat(recv, dim=2)
    ok_rows = torch.equal(rows, full[h0:h1])
    # row slabs -> column slabs
    w0, w1 = plan.cols[rank]
    send = [rows[:, lo:hi].contiguous() for lo, hi in plan.cols]
    recv = [torch.empty(plan.h_count(j), w1 - w0, Dp) for j in range(world)]
    comm.exchange([send], [recv])
    cols = torch.cat(recv, dim=0)
    ok_cols = torch.equal(cols, full[:, w0:w1])
    # column slabs -> d-slabs
    send = [cols[:, :, 4 * lo:4 * hi].contiguous() for lo, hi in plan.granules]
    recv = [torch.empty(H, plan.w_count(j), 4 * (g1 - g0)) for j in range(world)]
    comm.exchange([send], [recv])
    back = torch.cat(recv, dim=1)
    ok_back = torch.equal(back, slab)
    # WTA: per-slab (first minimum, its cost) -> all-gather -> first strict minimum in slab order
    cost = ((_volume(H, W, D) * 7919) % 13)                                     # many ties across slabs
    b, c = plan.d_base(rank), plan.d_count(rank)
    local = cost[:, :, b:b + c]
    part = torch.stack([local.argmin(dim=2).float() + b, local.min(dim=2).values])
    gathered = comm.all_gather([part])[0]                                       # [world][2][H][W]
    best = torch.full((H, W), float("inf")); idx = torch.full((H, W), -1.0)
    for s in range(world):
        take = gathered[s, 1] < best
        best = torch.where(take, gathered[s, 1], best); idx = torch.where(take, gathered[s, 0], idx)
    first_min = torch.from_numpy(np.argmin(cost.numpy(), axis=2).astype(np.float32))
    ok_wta = torch.equal(idx, first_min)
    # slab seam: the cells d*-1, d*, d*+1 summed over the ranks (each rank contributes what it owns)
    trip = torch.zeros(3, H, W)
    for k, off in enumerate((-1, 0, 1)):
        d = (idx + off).long()
        mine = (d >= b) & (d < b + c)
        trip[k] = torch.where(mine, cost.gather(2, d.clamp(0, D - 1)[..., None])[..., 0], torch.zeros(()))
    comm.all_reduce_sum([trip])
    want = torch.zeros(3, H, W)
    for k, off in enumerate((-1, 0, 1)):
        d = (idx + off).long()
        inside = (d >= 0) & (d < D)
        want[k] = torch.where(inside, cost.gather(2, d.clamp(0, D - 1)[..., None])[..., 0], torch.zeros(()))
    ok_seam = torch.equal(trip, want)
    res = torch.tensor([ok_rows, ok_cols, ok_back, ok_wta, ok_seam], dtype=torch.int32)
    dist.all_reduce(res, op=dist.ReduceOp.MIN)
    if rank == 0:
        np.save(os.path.join(out_dir, "ok.npy"), res.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,H,W,D", [(2, 9, 14, 22), (3, 7, 11, 31), (2, 4, 6, 8)])
def test_slab_repartitions_round_trip_over_gloo(tmp_path, world, H, W, D):
    mp.spawn(_worker, args=(world, _free_port(), H, W, D, str(tmp_path)), nprocs=world, join=True)
    ok = np.load(tmp_path / "ok.npy")
    assert ok.tolist() == [1, 1, 1, 1, 1], ok


def test_slab_plan_tiles_everything():
    sys.path.insert(0, ROOT)
    import importlib
    pkg = importlib.import_module("mc-cnn-python_b200")
    for (H, W, D) in ((2000, 3000, 400), (1024, 1024, 192), (9, 14, 22), (8, 8, 29)):
        for world in (1, 2, 3, 4, 8):
            if (D + 3) // 4 < world:
                continue
            p = pkg.SlabPlan(H, W, D, world)
            assert [p.d_base(r) for r in range(world)] == sorted(p.d_base(r) for r in range(world))
            assert all(p.d_base(r) % 4 == 0 and p.d_count(r) >= 1 for r in range(world))
            assert sum(p.d_count(r) for r in range(world)) == D
            assert p.d_base(0) == 0 and p.d_base(world - 1) + p.d_count(world - 1) == D
            assert sum(p.h_count(r) for r in range(world)) == H and sum(p.w_count(r) for r in range(world)) == W
            assert max(p.g_count(r) for r in range(world)) - min(p.g_count(r) for r in range(world)) <= 1
            assert [p.owner_of_disparity(d) for d in (0, D - 1)] == [0, world - 1]
    with pytest.raises(AssertionError):
        pkg.SlabPlan(100, 100, 8, 4)                      # two granules cannot feed four ranks


@pytest.mark.parametrize("world,H,W,D", [(2, 9, 14, 22), (3, 7, 11, 31), (4, 8, 9, 40)])
def test_local_comm_emulation_moves_the_same_blocks(world, H, W, D):
    """LocalComm (all ranks in one process: what the GPU tests drive the partition with) has the semantics of
    DistComm: exchange delivers sends[i][j] to recvs[j][i], all_gather stacks in rank order, all_reduce sums."""
    sys.path.insert(0, ROOT)
    import importlib
    pkg = importlib.import_module("mc-cnn-python_b200")
    plan = pkg.SlabPlan(H, W, D, world)
    comm = pkg.LocalComm(world)
    Dp = 4 * plan.G
    full = torch.zeros(H, W, Dp)
    full[:, :, :D] = _volume(H, W, D)
    slabs = [full[:, :, 4 * lo:4 * hi].contiguous() for lo, hi in plan.granules]
    sends = [[slabs[i][lo:hi].contiguous() for lo, hi in plan.rows] for i in range(world)]
    recvs = [[torch.empty(plan.h_count(j), W, 4 * plan.g_count(i)) for i in range(world)] for j in range(world)]
    comm.exchange(sends, recvs)
    for j, (lo, hi) in enumerate(plan.rows):
        assert torch.equal(torch.cat(recvs[j], dim=2), full[lo:hi])
    parts = [torch.full((2, 3), float(r)) for r in range(world)]
    for g in comm.all_gather(parts):
        assert g.shape == (world, 2, 3) and all(float(g[r, 0, 0]) == r for r in range(world))
    comm.all_reduce_sum(parts)
    assert all(torch.equal(p, torch.full((2, 3), float(sum(range(world))))) for p in parts)
