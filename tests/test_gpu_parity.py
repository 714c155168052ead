"""GPU: the CUDA path, called through the reference-named Python surface (which calls the C ABI),
against (a) the golden vectors produced by the reference's own NumPy code and (b) the C oracle on
seeded inputs.  Bit-exact (array_equal) for CBCA, SGM, WTA, interpolation, sub-pixel, median and
bilateral; scale-relative 1e-4 (north_star) for the cost volume; 2e-5 absolute on unit-norm
features for the CNN."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

DIRS = [(0, 1), (0, -1), (-1, 0), (1, 0)]
COST_RTOL = 1e-4          # BASELINE.json north_star: "fp32 cost volume within 1e-4 relative" (of the volume's scale)
FEAT_ATOL = 2e-5          # unit-norm features, fp32 accumulation over K = 576 x 5 layers


def eq(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


def synth_images(seed, H, W, levels, shift):
    """Blurred-noise image quantised to `levels` grey levels, normalised as match.py:118-123; right = shifted left."""
    rng = np.random.default_rng(seed)
    base = rng.random((H + 8, W + 8)).astype(np.float32)
    k = np.ones(5, np.float32) / 5
    for ax in (0, 1):
        base = np.apply_along_axis(lambda v: np.convolve(v, k, mode="same"), ax, base)
    base = base[4:-4, 4:-4]
    q = np.floor((base - base.min()) / (np.ptp(base) + 1e-9) * levels).astype(np.float32)
    qr = np.roll(q, -shift, axis=1)
    li = ((q - q.mean()) / q.std())[..., None].astype(np.float32)
    ri = ((qr - qr.mean()) / qr.std())[..., None].astype(np.float32)
    return li, ri


def unit_features(seed, H, W):
    rng = np.random.default_rng(seed)
    f = rng.standard_normal((2, H, W, 64)).astype(np.float32)
    f /= np.linalg.norm(f, axis=-1, keepdims=True)
    return f[0], f[1]


# ------------------------------------------------------------------------------------------ layout
@pytest.mark.parametrize("D,H,W", [(1, 3, 5), (11, 7, 33), (32, 16, 40), (70, 9, 65)])
def test_layout_round_trip(pf, D, H, W):
    import torch
    x = torch.randn(D, H, W, device="cuda")
    hwd, d, h, w = pf._as_hwd(x)
    assert (d, h, w) == (D, H, W) and hwd.shape == (H, W, (D + 3) // 4 * 4)
    assert torch.equal(hwd[:, :, :D].permute(2, 0, 1), x)
    assert torch.equal(pf._hwd_to_dhw(hwd, D), x)
    view = pf._hwd_view(hwd, D)
    assert pf._is_hwd_view(view) and pf._as_hwd(view)[0].data_ptr() == hwd.data_ptr()


# ------------------------------------------------------------------------------------------ features
def test_features_vs_golden_and_oracle(pf, pkg, oracle, features_golden):
    g = features_golden
    ws, bs = pf.glorot_uniform_weights(seed=int(g["glorot_seed"]))
    img = g["image"]
    fl, fr = pf.compute_features(img[..., None], img[::-1].copy()[..., None], 11, 11, (ws, bs))
    assert fl.shape == g["features_glorot"].shape and fl.dtype == np.float32
    np.testing.assert_allclose(fl, g["features_glorot"], atol=FEAT_ATOL, rtol=0)
    np.testing.assert_allclose(fr, oracle.net_forward(img[::-1].copy(), ws, bs), atol=FEAT_ATOL, rtol=0)
    np.testing.assert_allclose(np.linalg.norm(fl, axis=-1), 1.0, atol=1e-5)
    # ragged tile sizes + a different depth (patch 7 -> 3 layers)
    rng = np.random.default_rng(5)
    img2 = rng.standard_normal((37, 51)).astype(np.float32)
    f2, _ = pf.compute_features(img2[..., None], img2[..., None], 7, 7, (ws[:3], bs[:3]))
    np.testing.assert_allclose(f2, oracle.net_forward(img2, ws[:3], bs[:3]), atol=FEAT_ATOL, rtol=0)
    # model.NET applies no padding (VALID): features of the pre-padded image equal compute_features
    padded = np.pad(img, 5)[None, :, :, None]
    net = pkg.NET(padded)
    net.set_weights(ws, bs)
    assert net.features.shape == (1,) + fl.shape
    assert eq(net.features[0], fl)


def test_features_with_the_shipped_checkpoint(pf, oracle, features_golden, checkpoint_golden):
    """pf:32, :43: the reference restores model_epoch2000.ckpt; its ten conv tensors are a committed fixture, so the
    CUDA net runs the REAL weights here: against the float64 torch restatement (golden) and the C oracle."""
    g = features_golden
    ws, bs = checkpoint_golden
    img = g["image"]
    fl, fr = pf.compute_features(img[..., None], img[:, ::-1].copy()[..., None], 11, 11, (ws, bs))
    np.testing.assert_allclose(fl, g["features"], atol=FEAT_ATOL, rtol=0)
    np.testing.assert_allclose(fr, oracle.net_forward(img[:, ::-1].copy(), ws, bs), atol=FEAT_ATOL, rtol=0)
    np.testing.assert_allclose(np.linalg.norm(fl, axis=-1), 1.0, atol=1e-5)
    # a larger, ragged image (several tiles per row band) with the real weights, against the oracle
    rng = np.random.default_rng(8)
    big = rng.standard_normal((70, 300)).astype(np.float32)
    fb, _ = pf.compute_features(big[..., None], big[..., None], 11, 11, (ws, bs))
    np.testing.assert_allclose(fb, oracle.net_forward(big, ws, bs), atol=FEAT_ATOL, rtol=0)


def test_features_with_prepared_weights_are_the_bits_of_the_plain_call(pf):
    """mccnn_features_prepared (weights split once per network, what the Python surface uses) == mccnn_features."""
    import torch
    ffi = pf._ffi
    H, W, pad = 37, 150, 5
    img = torch.randn((H, W), device="cuda")
    dw = pf.resolve_weights(None, num_layers=5)
    a = pf.net_forward(img, dw, pad)
    b = torch.empty_like(a)
    nb = int(ffi.lib().mccnn_features_scratch_bytes(H, W, pad, 5))
    scratch = torch.empty((nb + 3) // 4, dtype=torch.float32, device="cuda")
    ffi.call("mccnn_features", ffi.ptr(img), H, W, pad, 5, dw.w_table, dw.b_table, ffi.ptr(b), ffi.ptr(scratch), ffi.stream_ptr())
    assert torch.equal(a, b)


def test_features_outside_the_fp16_range_take_the_tf32_kernel(pf, oracle, checkpoint_golden):
    """The tensor-core layers use FP16 hi/lo operands unless a layer's input or weights exceed fp16's range (65504): then the
    device-side gate sends that layer to the TF32 kernel.  Activations of 1e6 (first-layer weights scaled), weights of 1e5
    (third layer scaled), and both: the normalised features still match the float32 C oracle."""
    ws, bs = checkpoint_golden
    rng = np.random.default_rng(3)
    img = rng.standard_normal((40, 150)).astype(np.float32)
    for name, scale in (("activations", {0: 1e6}), ("weights", {2: 1e5}), ("both", {0: 3e5, 3: 2e5})):
        w2 = [w * np.float32(scale.get(i, 1.0)) for i, w in enumerate(ws)]
        b2 = [b * np.float32(np.prod([scale.get(j, 1.0) for j in range(i + 1)])) for i, b in enumerate(bs)]
        f, _ = pf.compute_features(img[..., None], img[..., None], 11, 11, (w2, b2))
        ref = oracle.net_forward(img, w2, b2)
        assert np.isfinite(f).all(), name
        np.testing.assert_allclose(f, ref, atol=FEAT_ATOL, rtol=0, err_msg=name)


def test_features_of_a_row_band_are_the_bits_of_the_whole_image(pf):
    """The net is local (11x11 receptive field), so rows [lo, hi) of the features need image rows [lo-5, hi+5) only --
    what the row-band feature exchange of the slab partition relies on.  The tensor-core layers must also give the SAME
    BITS for a band as for the whole image: every tile accumulates the two input-channel blocks in the same order
    wherever it falls in a CTA's sequence (round 1 alternated the order per tile ordinal: 1e-6 differences at sizes
    with many tiles per CTA, found by the slab parity leg of bench.py)."""
    import torch
    rng = np.random.default_rng(12)
    H, W = 700, 1500
    img = torch.from_numpy(rng.standard_normal((H, W)).astype(np.float32)).cuda()
    ws, bs = pf.glorot_uniform_weights(seed=0)
    full, _ = pf.compute_features(img[..., None], img[..., None], 11, 11, (ws, bs))
    for lo, hi in ((0, 130), (129, 402), (350, 700), (333, 334)):
        a, b = max(lo - 5, 0), min(hi + 5, H)
        band, _ = pf.compute_features(img[a:b, :, None], img[a:b, :, None], 11, 11, (ws, bs))
        # (rows of the band whose 5-row halo was cut by the band's own zero padding are not comparable)
        top, bot = (5 if a > 0 else 0), (5 if b < H else 0)
        assert torch.equal(band[top:band.shape[0] - bot], full[a + top:b - bot]), (lo, hi)


# ------------------------------------------------------------------------------------------ cost volume
def check_cost(got, ref):
    scale = float(np.abs(ref).max())
    assert got.shape == ref.shape and got.dtype == np.float32
    np.testing.assert_allclose(got, ref, atol=COST_RTOL * scale, rtol=0)


def test_cost_volume_vs_golden(pf, pipeline_golden):
    g = pipeline_golden
    L, R = pf.compute_cost_volume(g["fl"], g["fr"], int(g["ndisp"]))
    check_cost(L, g["cv_L"])
    check_cost(R, g["cv_R"])


@pytest.mark.parametrize("H,W,D", [(3, 4, 2), (5, 70, 33), (4, 300, 70), (3, 260, 192), (2, 515, 400)])
def test_cost_volume_vs_oracle(pf, oracle, H, W, D):
    fl, fr = unit_features(H * W + D, H, W)
    L, R = pf.compute_cost_volume(fl, fr, D)
    Lo, Ro = oracle.compute_cost_volume(fl, fr, D)
    check_cost(L, Lo)
    check_cost(R, Ro)
    # tensor in -> tensor out, logical [D,H,W] shape
    import torch
    Lt, Rt = pf.compute_cost_volume(torch.from_numpy(fl).cuda(), torch.from_numpy(fr).cuda(), D)
    assert tuple(Lt.shape) == (D, H, W) and Lt.is_cuda
    assert eq(Lt.cpu().numpy(), L) and eq(Rt.cpu().numpy(), R)


def test_cost_volume_rejects_what_the_reference_cannot_slice(pf):
    fl, fr = unit_features(0, 3, 9)
    with pytest.raises(AssertionError):
        pf.compute_cost_volume(fl, fr, 8)          # W < ndisp + 2


# ------------------------------------------------------------------------------------------ CBCA
def test_cross_regions_vs_golden(pf, oracle, pipeline_golden):
    g = pipeline_golden
    region, num = pf.compute_cross_region(g["left_image"], 0.02, 14)
    assert eq(num, g["region_num_left"]) and num.dtype == np.int32
    assert region.shape[2] == 28 * 28 and region.dtype == np.int32
    assert eq(region[:4], g["region_left_rows0_4"].astype(np.int32))
    arms, count = pf.cross_arms(g["right_image"], 0.02, 14)
    ao, co = oracle.cross_arms(g["right_image"], 0.02, 14)
    assert eq(arms.cpu().numpy(), ao) and eq(count.cpu().numpy(), co)


def test_flat_image_regions_and_aggregation_vs_golden(pf, flat_golden, monkeypatch):
    """Worst case of pf:585-599 (arms at the 13-pixel limit, regions up to 729): counts, explicit list and the
    aggregation against vectors from the reference's own code -- exact mode bit for bit, the separable mode within
    the re-association tolerance."""
    g = flat_golden
    region, num = pf.compute_cross_region(g["left_image"], 0.02, 14)
    assert num.max() == 729 and eq(num, g["region_num_left"])
    assert eq(region[13:15], g["region_left_rows13_15"].astype(np.int32))
    _, numr = pf.compute_cross_region(g["right_image"], 0.02, 14)
    assert eq(numr, g["region_num_right"])
    for iters in (1, 2, 5):
        monkeypatch.setattr(pf, "CBCA_MODE", pf.CBCA_EXACT)
        L, R = pf.cost_volume_aggregation(g["left_image"], g["right_image"], g["cv_L"], g["cv_R"], 0.02, 14, iters)
        assert eq(L, g["cbca%d_L" % iters]) and eq(R, g["cbca%d_R" % iters]), iters
        monkeypatch.setattr(pf, "CBCA_MODE", pf.CBCA_SEPARABLE)
        Ls, Rs = pf.cost_volume_aggregation(g["left_image"], g["right_image"], g["cv_L"], g["cv_R"], 0.02, 14, iters)
        scale = float(np.abs(g["cv_L"]).max())
        np.testing.assert_allclose(Ls, g["cbca%d_L" % iters], atol=CBCA_SEP_RTOL * scale, rtol=0)
        np.testing.assert_allclose(Rs, g["cbca%d_R" % iters], atol=CBCA_SEP_RTOL * scale, rtol=0)


@pytest.fixture
def exact_cbca(pf, monkeypatch):
    """Select the flat-running-sum kernel (bit-identical to the reference) for this test."""
    monkeypatch.setattr(pf, "CBCA_MODE", pf.CBCA_EXACT)


CBCA_SEP_RTOL = 2e-6      # separable mode: only the association of the float32 sum differs (scale-relative)


def test_cbca_bit_exact_vs_golden(pf, pipeline_golden, exact_cbca):
    g = pipeline_golden
    L, R = pf.cost_volume_aggregation(g["left_image"], g["right_image"], g["cv_L"], g["cv_R"], 0.02, 14, 2)
    assert eq(L, g["cbca1_L"]) and eq(R, g["cbca1_R"])
    L, R = pf.cost_volume_aggregation(g["left_image"], g["right_image"], g["sgm_L"], g["sgm_R"], 0.02, 14, 16)
    assert eq(L, g["cbca2_L"]) and eq(R, g["cbca2_R"])


@pytest.mark.parametrize("H,W,D,levels,iters", [(40, 90, 70, 4, 3), (33, 47, 192, 30, 1), (21, 35, 5, 2, 4),
                                                (6, 7, 2, 1, 2)])
def test_cbca_bit_exact_vs_oracle(pf, oracle, H, W, D, levels, iters, exact_cbca):
    li, ri = synth_images(H + W, H, W, levels, 2)
    rng = np.random.default_rng(D)
    L = rng.standard_normal((D, H, W)).astype(np.float32)
    R = rng.standard_normal((D, H, W)).astype(np.float32)
    Lc, Rc = L.copy(), R.copy()
    Lg, Rg = pf.cost_volume_aggregation(li, ri, L, R, 0.02, 14, iters)
    assert eq(L, Lc) and eq(R, Rc)                       # inputs untouched (pf:119)
    Lo, Ro = oracle.cost_volume_aggregation(li, ri, L, R, 0.02, 14, iters)
    assert eq(Lg, Lo) and eq(Rg, Ro)
    # zero rounds is the identity
    L0, _ = pf.cost_volume_aggregation(li, ri, L, R, 0.02, 14, 0)
    assert eq(L0, L)


def test_cbca_separable_vs_golden_and_oracle(pf, oracle, pipeline_golden):
    """Default (separable) mode: same region, same order within rows and along the spine, row sums
    formed first -> equal to the reference up to float32 re-association."""
    assert pf.CBCA_MODE == pf.CBCA_AUTO           # chained rounds or two passes per round, chosen per image; same bits
    g = pipeline_golden

    def close(a, b):
        scale = float(np.abs(b).max())
        np.testing.assert_allclose(a, b, atol=CBCA_SEP_RTOL * scale, rtol=0)

    L, R = pf.cost_volume_aggregation(g["left_image"], g["right_image"], g["cv_L"], g["cv_R"], 0.02, 14, 2)
    close(L, g["cbca1_L"]); close(R, g["cbca1_R"])
    L, R = pf.cost_volume_aggregation(g["left_image"], g["right_image"], g["sgm_L"], g["sgm_R"], 0.02, 14, 16)
    close(L, g["cbca2_L"]); close(R, g["cbca2_R"])
    for (H, W, D, levels, iters) in [(40, 90, 70, 4, 3), (33, 47, 192, 30, 1), (37, 29, 5, 1, 2), (50, 21, 33, 2, 5)]:
        li, ri = synth_images(H * W, H, W, levels, 2)
        rng = np.random.default_rng(D)
        Lv = rng.standard_normal((D, H, W)).astype(np.float32)
        Rv = rng.standard_normal((D, H, W)).astype(np.float32)
        Lg, Rg = pf.cost_volume_aggregation(li, ri, Lv, Rv, 0.02, 14, iters)
        Lo, Ro = oracle.cost_volume_aggregation(li, ri, Lv, Rv, 0.02, 14, iters)
        close(Lg, Lo); close(Rg, Ro)
    # integer-valued costs: sums are exact in any order -> bit-identical even in separable mode
    Li = rng.integers(0, 64, (12, 30, 44)).astype(np.float32)
    li, ri = synth_images(5, 30, 44, 3, 1)
    Lg, _ = pf.cost_volume_aggregation(li, ri, Li, Li, 0.02, 14, 1)
    Lo, _ = oracle.cost_volume_aggregation(li, ri, Li, Li, 0.02, 14, 1)
    assert eq(Lg, Lo)
    # arms longer than match.py's 13 pixels
    Lg, _ = pf.cost_volume_aggregation(li, ri, Li, Li, 0.02, 20, 1)
    Lo, _ = oracle.cost_volume_aggregation(li, ri, Li, Li, 0.02, 20, 1)
    assert eq(Lg, Lo)


def test_cbca_separable_many_shapes_vs_oracle(pf, oracle):
    """Default mode against the C oracle (pinned to the reference) within the re-association tolerance: flat images
    (13-pixel arms everywhere), widths around the 4-pixel patch and 64-pixel lines, granule counts that are not a
    multiple of 16, images shorter than an arm, one to five rounds, distance thresholds other than match.py's 14."""
    cases = [(40, 90, 70, 4, 3, 14), (33, 47, 192, 30, 2, 14), (9, 29, 33, 1, 2, 14), (50, 21, 40, 2, 3, 14),
             (64, 64, 32, 1, 1, 14), (17, 200, 29, 3, 2, 14), (70, 40, 100, 2, 2, 14), (12, 129, 8, 1, 4, 14),
             (25, 300, 20, 1, 5, 14), (31, 191, 68, 2, 3, 20), (14, 140, 12, 1, 2, 40), (20, 65, 6, 1, 3, 3),
             (5, 64, 4, 2, 2, 1), (3, 128, 130, 1, 2, 14)]
    for (H, W, D, levels, iters, dist) in cases:
        li, ri = synth_images(H * W + D, H, W, levels, 2)
        rng = np.random.default_rng(D)
        Lv = rng.standard_normal((D, H, W)).astype(np.float32) * 50
        Rv = rng.standard_normal((D, H, W)).astype(np.float32) * 50
        Lt, Rt = pf.cost_volume_aggregation(li, ri, Lv, Rv, 0.02, dist, iters)
        Lo, Ro = oracle.cost_volume_aggregation(li, ri, Lv, Rv, 0.02, dist, iters)
        scale = float(np.abs(Lo).max())
        np.testing.assert_allclose(Lt, Lo, atol=CBCA_SEP_RTOL * scale, rtol=0, err_msg=str((H, W, D, levels, iters, dist)))
        np.testing.assert_allclose(Rt, Ro, atol=CBCA_SEP_RTOL * scale, rtol=0)


def test_cbca_chained_rounds_match_two_pass(pf, monkeypatch):
    """The default mode chains the rounds of a call (rows | column pass of round k + row pass of round k+1 in one
    kernel, out_k in shared memory only | cols).  It forms the same sums in the same order as two streaming passes per
    round: IDENTICAL bits.  Covers flat images (13-pixel arms: the far-halo path of every segment and the global walks),
    widths around the 62-pixel segment (ragged last segment, arms crossing segment borders, a staged halo pixel that
    does not exist), granule counts that are not a multiple of 16, images shorter than an arm, two to five rounds,
    distance thresholds other than match.py's 14."""
    cases = [(40, 90, 70, 4, 3, 14), (33, 47, 192, 30, 2, 14), (9, 29, 33, 1, 2, 14), (50, 21, 40, 2, 3, 14),
             (64, 64, 32, 1, 2, 14), (17, 200, 29, 3, 2, 14), (70, 40, 100, 2, 2, 14), (12, 129, 8, 1, 4, 14),
             (25, 300, 20, 1, 5, 14), (31, 191, 68, 2, 3, 20), (14, 140, 12, 1, 2, 40), (20, 65, 6, 1, 3, 3),
             (5, 64, 4, 2, 2, 1), (3, 128, 130, 1, 2, 14), (30, 62, 16, 1, 3, 14), (30, 63, 16, 1, 3, 14),
             (30, 125, 16, 2, 3, 14), (300, 700, 48, 6, 4, 14)]
    for (H, W, D, levels, iters, dist) in cases:
        li, ri = synth_images(H * W + D, H, W, levels, 2)
        rng = np.random.default_rng(D)
        Lv = rng.standard_normal((D, H, W)).astype(np.float32) * 50
        Rv = rng.standard_normal((D, H, W)).astype(np.float32) * 50
        monkeypatch.setattr(pf, "CBCA_MODE", pf.CBCA_SEPARABLE_TWO_PASS)
        Ls, Rs = pf.cost_volume_aggregation(li, ri, Lv, Rv, 0.02, dist, iters)
        monkeypatch.setattr(pf, "CBCA_MODE", pf.CBCA_SEPARABLE)
        Lt, Rt = pf.cost_volume_aggregation(li, ri, Lv, Rv, 0.02, dist, iters)
        assert eq(Ls, Lt) and eq(Rs, Rt), (H, W, D, levels, iters, dist)


def test_cbca_settled_pixels_keep_the_bits_of_the_two_pass_form(pf, monkeypatch):
    """From the third pass of a chained call on, pixels without arms are skipped (both ping-pong buffers hold their value).
    Signed zeros, infinities and NaNs in such pixels, many rounds, every granules-per-thread shape: same bits as two
    passes per round, which never skip anything."""
    for (H, W, D, levels, iters) in [(40, 90, 192, 40, 9), (33, 70, 256, 25, 7), (24, 64, 20, 60, 16), (16, 50, 400, 30, 5)]:
        li, ri = synth_images(7 * H + D, H, W, levels, 2)
        rng = np.random.default_rng(H)
        Lv = (rng.integers(-3, 4, (D, H, W)) * 0.5).astype(np.float32)          # exact zeros among the values
        Lv[rng.random(Lv.shape) < 0.05] *= -0.0                                # ... of both signs
        Lv[:, 3, 5] = np.inf
        Lv[:, 7, 11] = np.nan
        Rv = -Lv
        arms, _ = pf.cross_arms(li, 0.02, 14)
        assert float((arms.reshape(-1, 4).sum(1) == 0).float().mean()) > 0.1, "the case needs pixels without arms"
        monkeypatch.setattr(pf, "CBCA_MODE", pf.CBCA_SEPARABLE_TWO_PASS)
        Ls, Rs = pf.cost_volume_aggregation(li, ri, Lv, Rv, 0.02, 14, iters)
        monkeypatch.setattr(pf, "CBCA_MODE", pf.CBCA_SEPARABLE)
        Lt, Rt = pf.cost_volume_aggregation(li, ri, Lv, Rv, 0.02, 14, iters)
        bits = lambda a: np.ascontiguousarray(a).view(np.uint32)
        assert eq(bits(Ls), bits(Lt)) and eq(bits(Rs), bits(Rt)), (H, W, D, iters)


def test_cbca_auto_mode_follows_the_vertical_arms(pf, pkg):
    """The host-side default picks the chained rounds for natural images and two passes per round when the vertical arms
    are long (piece-wise constant images); either way the bits are the same."""
    import torch
    from bench import synth_pair, flat_pair
    H, W, D = 96, 256, 24
    for maker, want in ((synth_pair, pf.CBCA_SEPARABLE), (flat_pair, pf.CBCA_SEPARABLE_TWO_PASS)):
        li, ri = maker(H, W, 9, seed=2)
        arms, _ = pf.cross_arms(li, 0.02, 14)
        assert pf.cbca_auto_mode(arms) == want
        m = pkg.StereoMatcher(H, W, D)
        m.set_images(li, ri)
        d_auto = m.run().clone()
        assert m.cbca_modes[0] == want
        for mode in (pf.CBCA_SEPARABLE, pf.CBCA_SEPARABLE_TWO_PASS):
            e = pkg.StereoMatcher(H, W, D, cbca_mode=mode)
            e.set_images(li, ri)
            assert torch.equal(e.run(), d_auto), mode


def test_cbca_every_mode_with_long_arms(pf, oracle, monkeypatch):
    """distance_threshold 20 (arms up to 19 pixels, longer than match.py's 13) on a flat image, every mode, integer
    costs: exact sums, so every mode must equal the oracle bit for bit after one round."""
    rng = np.random.default_rng(3)
    Li = rng.integers(0, 64, (12, 30, 90)).astype(np.float32)
    li, ri = synth_images(5, 30, 90, 1, 1)
    Lo, _ = oracle.cost_volume_aggregation(li, ri, Li, Li, 0.02, 20, 1)
    for mode in (pf.CBCA_SEPARABLE, pf.CBCA_EXACT, pf.CBCA_SEPARABLE_TWO_PASS):
        monkeypatch.setattr(pf, "CBCA_MODE", mode)
        Lg, _ = pf.cost_volume_aggregation(li, ri, Li, Li, 0.02, 20, 1)
        assert eq(Lg, Lo), mode
        L2, _ = pf.cost_volume_aggregation(li, ri, Li, Li, 0.02, 20, 2)
        Lo2, _ = oracle.cost_volume_aggregation(li, ri, Li, Li, 0.02, 20, 2)
        np.testing.assert_allclose(L2, Lo2, atol=CBCA_SEP_RTOL * 64, rtol=0)


def test_cbca_plane_constant_is_fixed_point(pf):
    """Property (any size): a volume that is constant per disparity plane with small-integer values is a
    fixed point of region averaging (exact sums, exact division)."""
    import torch
    H, W, D = 96, 160, 48
    li, ri = synth_images(9, H, W, 3, 1)
    vol = torch.arange(D, dtype=torch.float32, device="cuda")[:, None, None].expand(D, H, W).contiguous()
    L, R = pf.cost_volume_aggregation(li, ri, vol, vol, 0.02, 14, 3)
    assert torch.equal(L, vol) and torch.equal(R, vol)


# ------------------------------------------------------------------------------------------ SGM
def test_sgm_single_passes_bit_exact_and_in_place(pf, pipeline_golden):
    g = pipeline_golden
    for r in DIRS:
        p1 = 2.3 if r[0] == 0 else 2.3 / 1.5
        for ch, src in (("L", g["cbca1_L"]), ("R", g["cbca1_R"])):
            x = src.copy()
            y = pf.semi_global_matching(g["left_image"], g["right_image"], x, r, p1, 55.9, 4, 8, 0.08, ch)
            assert y is x                                   # reference aliasing (pf:544)
            assert eq(x, g["sgm_%s_%d_%d" % (ch, r[0], r[1])]), (r, ch)


def test_sgm_integer_costs_exact(pf, integer_golden):
    g = integer_golden
    for r in DIRS:
        for ch in "LR":
            x = g["R"].copy()
            pf.semi_global_matching(g["left_image"], g["right_image"], x, r, 2.0, 56.0, 4, 8, 0.08, ch)
            assert eq(x, g["sgm_int_%s_%d_%d" % (ch, r[0], r[1])]), (r, ch)


def test_sgm_average_is_four_chained_passes(pf, pipeline_golden):
    import torch
    g = pipeline_golden
    xl, xr = g["cbca1_L"].copy(), g["cbca1_R"].copy()
    L, R = pf.SGM_average(xl, xr, g["left_image"], g["right_image"], 2.3, 55.9, 4, 8, 0.08, 1.5)
    assert eq(L, g["sgm_L"]) and eq(R, g["sgm_R"])
    # the reference's passes run in place (pf:195-232, :544): the arrays handed in hold the result afterwards, and
    # what is returned is a different object (pf:232 builds a new array)
    assert eq(xl, g["sgm_L"]) and eq(xr, g["sgm_R"]) and L is not xl and R is not xr
    # tensor path: HWD views are updated in place and returned
    hl, D, _, _ = pf._as_hwd(torch.from_numpy(g["cbca1_L"]).cuda())
    hr, _, _, _ = pf._as_hwd(torch.from_numpy(g["cbca1_R"]).cuda())
    vl, vr = pf._hwd_view(hl, D), pf._hwd_view(hr, D)
    Lt, Rt = pf.SGM_average(vl, vr, g["left_image"], g["right_image"], 2.3, 55.9, 4, 8, 0.08, 1.5)
    assert Lt.data_ptr() == hl.data_ptr() and eq(Lt.cpu().numpy(), g["sgm_L"]) and eq(vr.cpu().numpy(), g["sgm_R"])


@pytest.mark.parametrize("H,W,D", [(9, 40, 2), (12, 45, 11), (10, 50, 128), (7, 210, 192), (6, 310, 300),
                                   (5, 420, 400)])
def test_sgm_vs_oracle_all_granule_shapes(pf, oracle, H, W, D):
    li, ri = synth_images(D, H, W, 12, 3)
    rng = np.random.default_rng(H * W)
    vol = (rng.standard_normal((D, H, W)) * 0.3).astype(np.float32)
    for r in DIRS:
        for ch in "LR":
            a, b = vol.copy(), vol.copy()
            pf.semi_global_matching(li, ri, a, r, 2.3, 55.9, 4, 8, 0.08, ch)
            oracle.semi_global_matching(li, ri, b, r, 2.3, 55.9, 4, 8, 0.08, ch)
            assert eq(a, b), (r, ch, float(np.abs(a - b).max()))
    a, b = pf.SGM_average(vol.copy(), vol.copy(), li, ri, 2.3, 55.9, 4, 8, 0.08, 1.5)
    ao, bo = oracle.SGM_average(vol.copy(), vol.copy(), li, ri, 2.3, 55.9, 4, 8, 0.08, 1.5)
    assert eq(a, ao) and eq(b, bo)


def test_sgm_argument_errors(pf, pipeline_golden):
    g = pipeline_golden
    x = g["cbca1_L"].copy()
    with pytest.raises(AssertionError):
        pf.semi_global_matching(g["left_image"], g["right_image"], x, (1, 1), 2.3, 55.9, 4, 8, 0.08, "L")   # pf:484
    with pytest.raises(AssertionError):
        pf.semi_global_matching(g["left_image"], g["right_image"], x, (0, 1), 2.3, 55.9, 4, 8, 0.08, "X")   # pf:479
    with pytest.raises(AssertionError):
        pf.semi_global_matching(g["left_image"], g["right_image"], x[:1].copy(), (0, 1), 2.3, 55.9, 4, 8, 0.08, "L")


def test_sgm_shift_property_full_size(pf):
    """Property at BASELINE config-3 width and depth: with integer costs and dyadic penalties the
    arithmetic is exact, and adding a constant K to the volume adds K to every output cell."""
    import torch
    H, W, D = 24, 1024, 192
    li, ri = synth_images(3, H, W, 64, 7)
    g = torch.Generator(device="cuda").manual_seed(1)
    base = torch.randint(0, 64, (H, W, D), generator=g, device="cuda").float()
    res = []
    for K in (0.0, 32.0):
        v = pf._hwd_view((base + K).contiguous(), D)
        L, _ = pf.SGM_average(v, v.clone(), li, ri, 2.0, 56.0, 4, 8, 0.08, 2.0)
        res.append(L.clone())
    assert torch.equal(res[0] + 32.0, res[1])


# ------------------------------------------------------------------------------------------ WTA + refinement
def test_wta_integer_costs_bit_identical(pf, integer_golden):
    g = integer_golden
    dl, dr = pf.disparity_prediction(g["L"], g["R"])
    assert dl.dtype == np.float32
    assert eq(dl, g["wta_L"]) and eq(dr, g["wta_R"])


@pytest.mark.parametrize("D,H,W", [(2, 5, 9), (11, 9, 31), (32, 16, 64), (100, 8, 50), (192, 16, 70), (400, 4, 33)])
def test_wta_first_minimum_rule(pf, D, H, W):
    rng = np.random.default_rng(D)
    vol = rng.integers(0, 6, (D, H, W)).astype(np.float32)      # many ties
    dl, dr = pf.disparity_prediction(vol, -vol)
    assert eq(dl, np.argmin(vol, axis=0).astype(np.float32))
    assert eq(dr, np.argmin(-vol, axis=0).astype(np.float32))


@pytest.mark.parametrize("D,H,W,levels,iters", [(2, 9, 21, 4, 1), (11, 13, 40, 6, 2), (70, 24, 75, 4, 3), (192, 20, 130, 3, 2), (400, 6, 45, 3, 2)])
def test_wta_folded_into_the_aggregation_is_the_separate_wta(pf, D, H, W, levels, iters):
    """mccnn_cbca_wta (the closing column pass of the aggregation takes the first minimum itself) returns the volume of
    mccnn_cbca and the map mccnn_wta makes of it, bit for bit: ties (quantised costs), -0.0 / +0.0, +inf and NaN cells,
    both separable schedules, with and without storing the volume."""
    import torch
    ffi = pf._ffi
    rng = np.random.default_rng(D + H)
    li, _ = synth_images(D, H, W, levels, 2)
    arms, count = pf.cross_arms(li, 0.02, 14)
    vol = (rng.integers(0, 5, (H, W, ffi.dpitch(D))) * 0.25).astype(np.float32)   # sums of quarters are exact: ties survive
    vol[rng.random(vol.shape) < 0.02] *= -0.0                                     # signed zeros among the ties
    vol[0, 0, :] = np.inf                                                         # a pixel with no finite cost at all
    vol[H - 1, W - 1, : max(1, D // 2)] = np.nan
    hwd = torch.from_numpy(vol).cuda()
    keys = torch.empty((H, W), dtype=torch.int64, device="cuda")
    for mode in (pf.CBCA_SEPARABLE, pf.CBCA_SEPARABLE_TWO_PASS):
        ref = pf._cbca_one(hwd, D, arms, count, iters, 14, mode=mode)
        want = pf._wta(ref, D)
        for store in (1, 0):
            out, scr = torch.empty_like(hwd), torch.empty_like(hwd)
            disp = torch.empty((H, W), dtype=torch.float32, device="cuda")
            ffi.call("mccnn_cbca_wta", ffi.ptr(hwd), ffi.ptr(out), ffi.ptr(scr), ffi.ptr(arms), ffi.ptr(count), D, H, W, iters, 14,
                     mode, store, ffi.ptr(keys), ffi.ptr(disp), ffi.stream_ptr())
            assert torch.equal(disp, want), (mode, store, int((disp != want).sum()))
            if store:
                assert eq(out[:, :, :D].cpu().numpy(), ref[:, :, :D].cpu().numpy())
    assert float(want[0, 0]) == -1.0


def test_wta_full_size_config2(pf):
    """BASELINE config 2 shape (512x512x128), integer-valued costs: indices identical to a first-minimum scan."""
    import torch
    D, H, W = 128, 512, 512
    g = torch.Generator(device="cuda").manual_seed(0)
    hwd = torch.randint(0, 256, (H, W, D), generator=g, device="cuda").float()
    vol = pf._hwd_view(hwd, D)
    dl, _ = pf.disparity_prediction(vol, vol)
    mn = vol.min(dim=0, keepdim=True).values
    first = ((vol == mn).cumsum(0) == 0).sum(0).float()
    assert torch.equal(dl, first)


def test_refinement_bit_exact_vs_golden(pf, pipeline_golden):
    g = pipeline_golden
    D = int(g["ndisp"])
    dl, dr = pf.disparity_prediction(g["cbca2_L"], g["cbca2_R"])
    assert eq(dl, g["wta_L"]) and eq(dr, g["wta_R"])
    d = pf.interpolation(dl, dr, D)
    assert eq(d, g["interp"])
    d = pf.subpixel_enhance(d, g["cbca2_L"])
    assert eq(d, g["subpixel"])
    d = pf.median_filter(d, 5, 5)
    assert eq(d, g["median"])
    d = pf.bilateral_filter(g["left_image"], d, 5, 5, 0, 6, 2)
    assert eq(d, g["bilateral"])


def test_refinement_integer_cases(pf, integer_golden):
    g = integer_golden
    D = g["L"].shape[0]
    assert eq(pf.interpolation(g["rand_dl"], g["rand_dr"], D), g["rand_interp"])
    assert eq(pf.subpixel_enhance(g["half_disp"], g["sub_vol"]), g["half_subpixel"])
    assert eq(pf.subpixel_enhance(g["rand_dl"], g["L"]), g["int_subpixel"])       # inf / NaN cells included
    assert eq(pf.median_filter(g["half_subpixel"], 5, 5), g["half_median"])
    assert eq(pf.bilateral_filter(g["left_image"], g["half_median"], 5, 5, 0, 6, 2), g["half_bilateral"])


def test_refinement_vs_oracle_random(pf, oracle):
    rng = np.random.default_rng(42)
    H, W, D = 45, 130, 40
    li, _ = synth_images(1, H, W, 200, 2)
    dl = rng.integers(0, D, (H, W)).astype(np.float32)
    dr = rng.integers(0, D, (H, W)).astype(np.float32)
    dl[:, 40:90] = 7.0
    dr[:, 30:85] = 7.0                                           # a consistent band so that label 0 exists
    out, lab = pf.interpolation(dl, dr, D, return_labels=True)
    oo, ol = oracle.interpolation(dl, dr, D, return_labels=True)
    assert eq(lab, ol) and eq(out, oo)
    assert set(np.unique(lab)) == {0, 1, 2}
    vol = rng.standard_normal((D, H, W)).astype(np.float32)
    half = (rng.integers(0, 2 * D - 1, (H, W)) / 2.0).astype(np.float32)
    sp = pf.subpixel_enhance(half, vol)
    assert eq(sp, oracle.subpixel_enhance(half, vol))
    for fh, fw in ((5, 5), (3, 7), (1, 1), (11, 11)):
        assert eq(pf.median_filter(sp, fh, fw), oracle.median_filter(sp, fh, fw))
        assert eq(pf.bilateral_filter(li, sp, fh, fw, 0, 6, 2), oracle.bilateral_filter(li, sp, fh, fw, 0, 6, 2))
    # 1x1 maps
    one = np.array([[1.0]], np.float32)
    assert pf.median_filter(one, 5, 5)[0, 0] == 1.0
    assert pf.bilateral_filter(np.zeros((1, 1, 1), np.float32), one, 5, 5, 0, 6, 2)[0, 0] == 1.0


# ------------------------------------------------------------------------------------------ whole pipeline
def test_pipeline_object_matches_stagewise_functions_and_oracle(pkg, pf, oracle, pipeline_golden):
    g = pipeline_golden
    D = int(g["ndisp"])
    H, W = g["left_image"].shape[:2]
    stages = tuple(s for s in pkg.pipeline.STAGES if s != "features")
    m = pkg.StereoMatcher(H, W, D, stages=stages)
    m.set_images(g["left_image"], g["right_image"])
    m.set_features(g["fl"], g["fr"])
    d = m.run().cpu().numpy()
    # stage-wise through the reference-named functions on the same inputs
    L, R = pf.compute_cost_volume(g["fl"], g["fr"], D)
    L, R = pf.cost_volume_aggregation(g["left_image"], g["right_image"], L, R, 0.02, 14, 2)
    L, R = pf.SGM_average(L, R, g["left_image"], g["right_image"], 2.3, 55.9, 4, 8, 0.08, 1.5)
    L, R = pf.cost_volume_aggregation(g["left_image"], g["right_image"], L, R, 0.02, 14, 16)
    assert eq(m.volume(0).cpu().numpy(), L) and eq(m.volume(1).cpu().numpy(), R)
    dl, dr = pf.disparity_prediction(L, R)
    e = pf.bilateral_filter(g["left_image"], pf.median_filter(pf.subpixel_enhance(pf.interpolation(dl, dr, D), L), 5, 5),
                            5, 5, 0, 6, 2)
    assert eq(d, e)
    # against the reference's end result: only the 64-term dot product is reassociated (~1e-7), so the final
    # volume agrees to scale-relative 1e-4 and the disparity map agrees except for rare near-tie flips
    scale = float(np.abs(g["cbca2_L"]).max())
    np.testing.assert_allclose(L, g["cbca2_L"], atol=1e-4 * scale, rtol=0)
    flips = int(np.sum(np.abs(d - g["bilateral"]) >= 1e-3))
    print("pipeline object vs reference golden: %d of %d pixels differ by >= 1e-3" % (flips, d.size))
    assert flips <= max(1, d.size // 1000), flips


def test_full_pipeline_recovers_known_shift(pkg):
    """128x128x32 (BASELINE config 1 shape), random-init network, right image = left shifted by 5 px:
    the final map is 5 away from the borders."""
    H, W, D, shift = 128, 128, 32, 5
    rng = np.random.default_rng(0)
    base = rng.random((H, W + shift)).astype(np.float32) * 255
    k = np.ones(3, np.float32) / 3
    base = np.apply_along_axis(lambda v: np.convolve(v, k, mode="same"), 1, base)
    left, right = np.floor(base[:, :W]), np.floor(base[:, shift:])       # right(x) = left(x + shift)
    li = ((left - left.mean()) / left.std())[..., None].astype(np.float32)
    ri = ((right - right.mean()) / right.std())[..., None].astype(np.float32)
    d = pkg.match_pair(li, ri, D)
    assert d.shape == (H, W) and d.dtype == np.float32
    inner = d[8:-8, 40:-8]
    assert np.mean(np.abs(inner - shift) < 0.5) > 0.95, float(np.mean(np.abs(inner - shift) < 0.5))


# ------------------------------------------------------------------------------------------ BASELINE configs
def test_config1_full_pipeline_stagewise_vs_oracle(pkg, pf, oracle):
    """BASELINE config 1 (the correctness gate): 128x128 pair, 32 disparities, random-init MC-CNN-fast features,
    the whole match.py pipeline.  Features against the CPU oracle; every later stage is fed the oracle's input for
    that stage (stage-wise gating keeps float near-ties from hiding real differences)."""
    H, W, D = 128, 128, 32
    li, ri = synth_images(11, H, W, 40, 3)
    ws, bs = pf.glorot_uniform_weights(seed=0)
    fl, fr = pf.compute_features(li, ri, 11, 11, (ws, bs))
    flo, fro = oracle.compute_features(li, ri, 11, 11, (ws, bs))
    assert np.abs(fl - flo).max() < FEAT_ATOL and np.abs(fr - fro).max() < FEAT_ATOL
    do, st = oracle.match_from_features(li, ri, flo, fro, D, return_stages=True)
    L, R = pf.compute_cost_volume(flo, fro, D)
    scale = float(np.abs(st["cost_volume"][0]).max())
    np.testing.assert_allclose(L, st["cost_volume"][0], atol=COST_RTOL * scale, rtol=0)
    np.testing.assert_allclose(R, st["cost_volume"][1], atol=COST_RTOL * scale, rtol=0)
    L, R = pf.cost_volume_aggregation(li, ri, *st["cost_volume"], 0.02, 14, 2)
    scale = float(np.abs(st["cbca1"][0]).max())
    np.testing.assert_allclose(L, st["cbca1"][0], atol=CBCA_SEP_RTOL * scale, rtol=0)
    L, R = pf.SGM_average(st["cbca1"][0].copy(), st["cbca1"][1].copy(), li, ri, 2.3, 55.9, 4, 8, 0.08, 1.5)
    assert eq(L, st["sgm"][0]) and eq(R, st["sgm"][1])                      # bit-exact
    L, R = pf.cost_volume_aggregation(li, ri, *st["sgm"], 0.02, 14, 16)
    scale = float(np.abs(st["cbca2"][0]).max())
    np.testing.assert_allclose(L, st["cbca2"][0], atol=16 * CBCA_SEP_RTOL * scale, rtol=0)
    dl, dr = pf.disparity_prediction(*st["cbca2"])
    assert eq(dl, st["wta"][0]) and eq(dr, st["wta"][1])                     # bit-identical indices
    d = pf.interpolation(*st["wta"], D)
    assert eq(d, st["interpolation"])
    d = pf.subpixel_enhance(st["interpolation"], st["cbca2"][0])
    assert eq(d, st["subpixel"])
    d = pf.median_filter(st["subpixel"], 5, 5)
    assert eq(d, st["median"])
    d = pf.bilateral_filter(li, st["median"], 5, 5, 0, 6, 2)
    np.testing.assert_allclose(d, st["bilateral"], rtol=2e-6, atol=1e-6)
    # end to end (everything on the GPU, CUDA features): the maps agree except for rare near-tie flips
    d2 = pkg.match_pair(li, ri, D, checkpoint=(ws, bs))
    flips = int(np.sum(np.abs(d2 - do) >= 1e-3))
    print("config 1 end to end: %d of %d pixels differ from the oracle's final map by >= 1e-3" % (flips, d2.size))
    assert flips <= d2.size // 1000, flips


def test_config2_cost_volume_and_wta_vs_oracle(pf, oracle):
    """BASELINE config 2: 512x512x128, cost volume + WTA only.  Cost volume within 1e-4 of the volume's scale; WTA of
    the oracle's volume bit-identical; WTA of our own volume agrees except where the two smallest costs are closer than
    the tolerance."""
    H, W, D = 512, 512, 128
    fl, fr = unit_features(7, H, W)
    L, R = pf.compute_cost_volume(fl, fr, D)
    Lo, Ro = oracle.compute_cost_volume(fl, fr, D)
    scale = float(np.abs(Lo).max())
    assert np.abs(L - Lo).max() <= COST_RTOL * scale and np.abs(R - Ro).max() <= COST_RTOL * scale
    dl, dr = pf.disparity_prediction(Lo, Ro)
    dlo, dro = oracle.disparity_prediction(Lo, Ro)
    assert eq(dl, dlo) and eq(dr, dro)
    dl2, _ = pf.disparity_prediction(L, R)
    assert np.mean(dl2 == dlo) > 0.999


# ------------------------------------------------------------------------------------------ match.py drop-in
def test_match_cli_writes_middlebury_outputs(pkg, tmp_path):
    """The match.py drop-in: list file + calib + images in, PFM / PGM / time file out (match.py:46-54, :182-184),
    identical to match_pair on the same pair; -s/-e window respected."""
    import importlib
    cv2 = pytest.importorskip("cv2")
    match = importlib.import_module("mc-cnn-python_b200.match")
    util = importlib.import_module("mc-cnn-python_b200.util")
    data = tmp_path / "data"
    H, W, D = 40, 72, 16
    rng = np.random.default_rng(5)
    paths = []
    for name in ("A", "B", "C"):
        d = data / name
        d.mkdir(parents=True)
        base = rng.integers(0, 256, (H, W + 4)).astype(np.uint8)
        base = cv2.GaussianBlur(base, (5, 5), 1.0)
        cv2.imwrite(str(d / "im0.png"), base[:, :W])
        cv2.imwrite(str(d / "im1.png"), base[:, 4:])
        (d / "calib.txt").write_text("cam0=[]\ncam1=[]\ndoffs=0\nbaseline=1\nwidth=%d\nheight=%d\nndisp=%d\n" % (W, H, D))
        paths.append(str(d / "im0.png"))
    lst = tmp_path / "list.txt"
    lst.write_text("\n".join(paths) + "\n")
    save = tmp_path / "out"
    save.mkdir()
    done = match.main(["--list_file", str(lst), "--data_dir", str(data), "--save_dir", str(save), "-t", "x",
                       "-s", "1", "-e", "2"])
    assert done == [1, 2]
    assert not (save / "submit_x" / "A").exists()
    for name in ("B", "C"):
        res = save / "submit_x" / name
        pfm = util.readPfm(str(res / "disp0MCCNN.pfm"))
        assert pfm.shape == (H, W) and float((res / "timeMCCNN.txt").read_text()) > 0
        assert (save / "submit_x_imgs" / name / "disp0MCCNN.pgm").read_bytes().startswith(b"P5")
        li = match.read_normalised(str(data / name / "im0.png"))
        ri = match.read_normalised(str(data / name / "im1.png"))
        ref = pkg.match_pair(li, ri, D)
        assert eq(pfm, ref)
