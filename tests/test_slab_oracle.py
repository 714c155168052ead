"""CPU: the facts the disparity-slab partition (mc-cnn-python_b200/slab.py) rests on, checked on the
oracle (the C restatement of the reference) without any CUDA code: cross-based aggregation is independent per
disparity plane, the horizontal SGM passes are independent per row, a per-slab first minimum combined in slab order is
the reference's first minimum, and the three sub-pixel cells summed over their owners are the cells themselves."""
import numpy as np
import pytest


def _images(seed, H, W, levels=6, shift=2):
    rng = np.random.default_rng(seed)
    base = rng.random((H, W + shift)).astype(np.float32)
    k = np.ones(3, np.float32) / 3
    for ax in (0, 1):
        base = np.apply_along_axis(lambda v: np.convolve(v, k, mode="same"), ax, base)
    q = np.floor((base - base.min()) / (np.ptp(base) + 1e-9) * levels).astype(np.float32)
    l, r = q[:, :W], q[:, shift:]
    norm = lambda a: ((a - a.mean()) / a.std())[..., None].astype(np.float32)
    return norm(l), norm(r)


@pytest.mark.parametrize("H,W,D,slabs", [(14, 22, 12, [(0, 4), (4, 12)]), (9, 30, 11, [(0, 4), (4, 8), (8, 11)])])
def test_aggregation_is_independent_per_disparity_plane(oracle, H, W, D, slabs):
    li, ri = _images(H * W, H, W)
    rng = np.random.default_rng(D)
    L = rng.standard_normal((D, H, W)).astype(np.float32)
    R = rng.standard_normal((D, H, W)).astype(np.float32)
    Lf, Rf = oracle.cost_volume_aggregation(li, ri, L, R, 0.02, 14, 3)
    for lo, hi in slabs:
        Ls, Rs = oracle.cost_volume_aggregation(li, ri, L[lo:hi].copy(), R[lo:hi].copy(), 0.02, 14, 3)
        assert np.array_equal(Ls, Lf[lo:hi]) and np.array_equal(Rs, Rf[lo:hi])


@pytest.mark.parametrize("r", [(0, 1), (0, -1)])
def test_horizontal_sgm_passes_are_independent_per_row(oracle, r):
    H, W, D = 11, 26, 9
    li, ri = _images(3, H, W)
    rng = np.random.default_rng(1)
    vol = rng.standard_normal((D, H, W)).astype(np.float32)
    for choice in "LR":
        whole = oracle.semi_global_matching(li, ri, vol.copy(), r, 2.3, 55.9, 4, 8, 0.08, choice)
        for lo, hi in ((0, 4), (4, 5), (5, 11)):
            part = oracle.semi_global_matching(li[lo:hi].copy(), ri[lo:hi].copy(), vol[:, lo:hi].copy(), r, 2.3, 55.9, 4, 8,
                                               0.08, choice)
            assert np.array_equal(part, whole[:, lo:hi]), (choice, lo, hi)


def test_slab_winners_and_seam_cells_reconstruct_the_reference(oracle):
    H, W, D = 10, 17, 23
    rng = np.random.default_rng(7)
    L = rng.integers(0, 9, (D, H, W)).astype(np.float32)          # many ties, also across slab borders
    R = rng.integers(0, 9, (D, H, W)).astype(np.float32)
    dl, dr = oracle.disparity_prediction(L, R)
    bounds = [(0, 8), (8, 16), (16, 23)]
    for vol, want in ((L, dl), (R, dr)):
        best = np.full((H, W), np.inf, np.float32); idx = np.full((H, W), -1, np.float32)
        for lo, hi in bounds:                                      # slab order, strict <: the lowest disparity wins ties
            local = oracle.wta_one(vol[lo:hi].copy()) + lo
            val = np.take_along_axis(vol, local.astype(np.int64)[None], 0)[0]
            take = val < best
            best = np.where(take, val, best); idx = np.where(take, local, idx)
        assert np.array_equal(idx, want)
    # sub-pixel: C[d-1], C[d], C[d+1] gathered from their owners (zeros elsewhere) and summed
    d = dl.copy()
    trip = np.zeros((3, H, W), np.float32)
    for lo, hi in bounds:
        for k, off in enumerate((-1, 0, 1)):
            dd = (d + off).astype(np.int64)
            mine = (dd >= lo) & (dd < hi)
            trip[k] += np.where(mine, np.take_along_axis(L, np.clip(dd, 0, D - 1)[None], 0)[0], 0).astype(np.float32)
    inside = (d - 1 >= 0) & (d + 1 < D)
    Cm, C, Cp = trip
    with np.errstate(divide="ignore", invalid="ignore"):
        mine_sub = np.where(inside, d - (Cp - Cm) / (np.float32(2.0) * ((Cp - np.float32(2.0) * C) + Cm)), d).astype(np.float32)
        want = oracle.subpixel_enhance(d, L)
    assert np.array_equal(mine_sub, want, equal_nan=True)
