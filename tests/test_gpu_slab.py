"""GPU: the disparity-slab partition of one big pair (BASELINE config C5) against the single-GPU
pipeline and the oracle.  All ranks of the partition run inside this process on one GPU (LocalComm): the same
SlabRank code, packing kernels and phase order as under torch.distributed, with device copies for the exchanges."""
import ctypes

import numpy as np
import pytest

from test_gpu_parity import eq, synth_images, unit_features, check_cost

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("H,W,D,world", [(5, 70, 33, 2), (4, 300, 70, 3), (4, 260, 192, 4), (6, 150, 19, 2)])
def test_cost_volume_slabs_tile_the_full_volume(pkg, pf, oracle, H, W, D, world):
    """Every slab equals the corresponding disparity planes of the full-D call (bit for bit: the dot product of a
    cell does not depend on where its tile sits) and of the oracle (1e-4 of scale)."""
    import torch
    ffi = pkg._ffi
    fl, fr = unit_features(H * W + D, H, W)
    L, R = pf.compute_cost_volume(fl, fr, D)
    Lo, Ro = oracle.compute_cost_volume(fl, fr, D)
    check_cost(L, Lo); check_cost(R, Ro)
    plan = pkg.SlabPlan(H, W, D, world)
    tl, tr = torch.from_numpy(fl).cuda(), torch.from_numpy(fr).cuda()
    for r in range(world):
        base, cnt, pitch = plan.d_base(r), plan.d_count(r), 4 * plan.g_count(r)
        sl = torch.full((H, W, pitch), 7.0, device="cuda"); sr = torch.full((H, W, pitch), 7.0, device="cuda")
        ffi.call("mccnn_cost_volume_slab", ffi.ptr(tl), ffi.ptr(tr), ffi.ptr(sl), ffi.ptr(sr), H, W, 64, D, base, cnt,
                 ffi.stream_ptr())
        got_l = sl[:, :, :cnt].permute(2, 0, 1).cpu().numpy()
        got_r = sr[:, :, :cnt].permute(2, 0, 1).cpu().numpy()
        assert eq(got_l, L[base:base + cnt]) and eq(got_r, R[base:base + cnt]), (r, base, cnt)
    assert sum(plan.d_count(r) for r in range(world)) == D


def test_copy3d_packs_and_unpacks_blocks(pkg):
    import torch
    ffi = pkg._ffi
    H, W, G = 7, 11, 5
    src = torch.randn(H, W, 4 * G, device="cuda")
    # columns [3, 9) of every row, packed
    dst = torch.zeros(H, 6, 4 * G, device="cuda")
    ffi.call("mccnn_copy3d", ctypes.c_void_p(src.data_ptr() + 16 * 3 * G), ffi.ptr(dst), H, 6, G, W * G, G, 6 * G, G,
             ffi.stream_ptr())
    assert torch.equal(dst, src[:, 3:9])
    # granules [1, 4) of every pixel, packed, and back into a zeroed volume
    pk = torch.zeros(H, W, 12, device="cuda")
    ffi.call("mccnn_copy3d", ctypes.c_void_p(src.data_ptr() + 16), ffi.ptr(pk), 1, H * W, 3, 0, G, 0, 3, ffi.stream_ptr())
    assert torch.equal(pk, src[:, :, 4:16])
    back = torch.zeros_like(src)
    ffi.call("mccnn_copy3d", ffi.ptr(pk), ctypes.c_void_p(back.data_ptr() + 16), 1, H * W, 3, 0, 3, 0, G, ffi.stream_ptr())
    assert torch.equal(back[:, :, 4:16], src[:, :, 4:16]) and float(back[:, :, :4].abs().sum()) == 0.0


@pytest.mark.parametrize("H,W,D,world", [(40, 96, 48, 2), (33, 75, 37, 3), (48, 130, 100, 4), (21, 64, 31, 8),
                                         (420, 1300, 200, 2), (300, 900, 96, 3)])
def test_slab_partition_equals_single_gpu_pipeline(pkg, H, W, D, world):
    """The partitioned pipeline returns the single-GPU pipeline's disparity map bit for bit, and so do its
    intermediate volumes (every cell is computed by the same kernel on the same operands, only elsewhere)."""
    import torch
    li, ri = synth_images(H + W + D, H, W, 40, 3)
    one = pkg.StereoMatcher(H, W, D)
    one.set_images(li, ri)
    want = one.run().clone()
    plan = pkg.SlabPlan(H, W, D, world)
    ranks = [pkg.SlabRank(plan, r) for r in range(world)]
    for rk in ranks:
        rk.set_images(li, ri)
    maps = pkg.run_slabs(ranks, pkg.LocalComm(world))
    torch.cuda.synchronize()
    full = one.final_volume                                   # [H][W][Dp] left / right after CBCA x 16
    for r, rk in enumerate(ranks):
        b, c = plan.d_base(r), plan.d_count(r)
        for v in range(2):
            assert torch.equal(rk.volA[v][:, :, :c], full[v][:, :, b:b + c]), (r, v)
    for m in maps:
        assert torch.equal(m, want)
    assert torch.equal(ranks[0].disp[0], one.disp[0]) and torch.equal(ranks[-1].disp[1], one.disp[1])


@pytest.mark.parametrize("H,W,D,world", [(40, 96, 48, 2), (33, 75, 37, 3), (48, 130, 100, 4), (21, 64, 31, 8)])
def test_slab_partition_over_peer_memory_equals_single_gpu_pipeline(pkg, H, W, D, world):
    """The same partition with the re-partitions fused into the kernels (strided peer copies, SGM passes that store
    into the next owner's slab): still the single-GPU map and volumes, bit for bit."""
    import torch
    li, ri = synth_images(H + W + D, H, W, 40, 3)
    one = pkg.StereoMatcher(H, W, D)
    one.set_images(li, ri)
    want = one.run().clone()
    plan = pkg.SlabPlan(H, W, D, world)
    comm = pkg.LocalComm(world)
    arenas = comm.make_arenas(6 * plan.region_floats())
    ranks = [pkg.SlabRank(plan, r, arena=arenas[r]) for r in range(world)]
    for rk in ranks:
        rk.set_images(li, ri)
    for _ in range(2):                                        # twice: the arenas are reused from pair to pair
        maps = pkg.run_slabs_p2p(ranks, comm)
    torch.cuda.synchronize()
    full = one.final_volume
    for r, rk in enumerate(ranks):
        b, c = plan.d_base(r), plan.d_count(r)
        for v in range(2):
            assert torch.equal(rk.volA[v][:, :, :c], full[v][:, :, b:b + c]), (r, v)
    for m in maps:
        assert torch.equal(m, want)


def test_slab_sgm_passes_match_whole_image_passes(pkg, pf):
    """Row slabs under the horizontal passes and column slabs under the vertical passes reproduce the whole-image
    passes (pf:195-208) exactly, including the other image's penalty look-up across the slab edge."""
    import torch
    ffi = pkg._ffi
    H, W, D = 19, 83, 40
    li, ri = synth_images(3, H, W, 30, 2)
    vol = torch.randn(2, H, W, D, device="cuda") * 3
    il, ir = pf._to_dev(li[..., 0]), pf._to_dev(ri[..., 0])
    flags = torch.empty((int(ffi.lib().mccnn_sgm_scratch_bytes(H, W, D)) + 3) // 4, dtype=torch.int32, device="cuda")
    args = (2.3, 55.9, 4.0, 8.0, 0.08, 1.5)
    ref = vol.clone()
    ffi.call("mccnn_sgm_average_pair", ffi.ptr(ref[0]), ffi.ptr(ref[1]), ffi.ptr(il), ffi.ptr(ir), ffi.ptr(flags), D, H, W,
             *args, ffi.stream_ptr())
    got = vol.clone()
    for lo, hi in ((0, 7), (7, 8), (8, 19)):                   # row slabs
        a, b = got[0, lo:hi].contiguous(), got[1, lo:hi].contiguous()
        ffi.call("mccnn_sgm_passes_slab", ffi.ptr(a), ffi.ptr(b), ctypes.c_void_p(il.data_ptr() + 4 * lo * W),
                 ctypes.c_void_p(ir.data_ptr() + 4 * lo * W), ffi.ptr(flags), D, hi - lo, W, 0, W, 0, *args, ffi.stream_ptr())
        got[0, lo:hi] = a; got[1, lo:hi] = b
    for lo, hi in ((0, 30), (30, 31), (31, 83)):               # column slabs
        a, b = got[0, :, lo:hi].contiguous(), got[1, :, lo:hi].contiguous()
        ffi.call("mccnn_sgm_passes_slab", ffi.ptr(a), ffi.ptr(b), ffi.ptr(il), ffi.ptr(ir), ffi.ptr(flags), D, H, W, lo,
                 hi - lo, 1, *args, ffi.stream_ptr())
        got[0, :, lo:hi] = a; got[1, :, lo:hi] = b
    assert torch.equal(got, ref)
