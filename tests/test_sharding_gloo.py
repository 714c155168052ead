"""CPU, world_size 2 over gloo: the image-pair sharding of the multi-GPU path (one process per GPU, pairs
partitioned in disjoint windows like the reference's -s/-e flags, no collective on the data path; only the final
[H,W] maps are gathered for the writer rank)."""
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


def _worker(rank, world, port, num_pairs, out_dir):
    sys.path.insert(0, ROOT)
    import importlib
    pkg = importlib.import_module("mc-cnn-python_b200")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    start, end = pkg.shard_window(num_pairs, rank, world)
    # stand-in for the per-pair result of this rank's window: a map whose value is the pair index
    H, W = 4, 6
    mine = torch.full((num_pairs, H, W), -1.0)
    for i in range(start, end):
        mine[i] = float(i)
    # gather of the final maps on the writer rank (rank 0); max works because unowned slots are -1
    dist.reduce(mine, dst=0, op=dist.ReduceOp.MAX)
    window = torch.tensor([start, end])
    windows = [torch.zeros(2, dtype=torch.long) for _ in range(world)]
    dist.all_gather(windows, window)
    if rank == 0:
        np.save(os.path.join(out_dir, "maps.npy"), mine.numpy())
        np.save(os.path.join(out_dir, "windows.npy"), torch.stack(windows).numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("num_pairs", [1, 2, 7, 8])
def test_pairs_are_partitioned_disjointly_and_gathered(tmp_path, num_pairs):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), num_pairs, str(tmp_path)), nprocs=world, join=True)
    maps = np.load(tmp_path / "maps.npy")
    windows = np.load(tmp_path / "windows.npy")
    assert windows[0, 0] == 0 and windows[-1, 1] == num_pairs
    assert all(windows[r, 1] == windows[r + 1, 0] for r in range(world - 1))          # contiguous, disjoint
    assert max(int(e - s) for s, e in windows) - min(int(e - s) for s, e in windows) <= 1   # balanced
    for i in range(num_pairs):
        assert np.all(maps[i] == i)                                                       # every pair done exactly once


def test_shard_window_matches_reference_flag_semantics():
    sys.path.insert(0, ROOT)
    import importlib
    pkg = importlib.import_module("mc-cnn-python_b200")
    for n in (0, 1, 5, 8, 15, 16, 17):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                s, e = pkg.shard_window(n, r, world)
                seen += list(range(s, e))                     # match.py:85-90 processes indices in [start, end)
            assert seen == list(range(n))
