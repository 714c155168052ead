"""CPU: the C oracle (oracle/mccnn_oracle.c) against the golden vectors that were produced by
running the REFERENCE's own process_functional.py (oracle/gen_golden.py).  Bit-exact for every
post-CNN stage -- this is what pins the oracle."""
import numpy as np

HP = dict(tau=0.02, dist=14)
DIRS = [(0, 1), (0, -1), (-1, 0), (1, 0)]


def eq(a, b):
    return np.array_equal(a, b, equal_nan=True)


def test_cost_volume_bit_exact(oracle, pipeline_golden):
    g = pipeline_golden
    L, R = oracle.compute_cost_volume(g["fl"], g["fr"], int(g["ndisp"]))
    assert L.dtype == np.float32 and L.shape == g["cv_L"].shape
    assert eq(L, g["cv_L"]) and eq(R, g["cv_R"])


def test_cross_region_matches_reference_lists(oracle, pipeline_golden):
    g = pipeline_golden
    region, num = oracle.compute_cross_region(g["left_image"], 0.02, 14)
    assert eq(num, g["region_num_left"])
    assert eq(region[:4], g["region_left_rows0_4"].astype(np.int32))
    assert region.shape[2] == 28 * 28 and region.dtype == np.int32


def test_cbca_bit_exact(oracle, pipeline_golden):
    g = pipeline_golden
    L, R = oracle.cost_volume_aggregation(g["left_image"], g["right_image"], g["cv_L"], g["cv_R"], 0.02, 14, 2)
    assert eq(L, g["cbca1_L"]) and eq(R, g["cbca1_R"])
    L, R = oracle.cost_volume_aggregation(g["left_image"], g["right_image"], g["sgm_L"], g["sgm_R"], 0.02, 14, 16)
    assert eq(L, g["cbca2_L"]) and eq(R, g["cbca2_R"])


def test_sgm_single_passes_bit_exact_and_in_place(oracle, pipeline_golden):
    g = pipeline_golden
    for r in DIRS:
        p1 = 2.3 if r[0] == 0 else 2.3 / 1.5
        for ch, src in (("L", g["cbca1_L"]), ("R", g["cbca1_R"])):
            x = src.copy()
            y = oracle.semi_global_matching(g["left_image"], g["right_image"], x, r, p1, 55.9, 4, 8, 0.08, ch)
            assert y is x                                   # reference aliasing (pf:544)
            assert eq(x, g["sgm_%s_%d_%d" % (ch, r[0], r[1])])


def test_sgm_average_is_four_chained_passes(oracle, pipeline_golden):
    g = pipeline_golden
    L, R = oracle.SGM_average(g["cbca1_L"].copy(), g["cbca1_R"].copy(), g["left_image"], g["right_image"],
                              2.3, 55.9, 4, 8, 0.08, 1.5)
    assert eq(L, g["sgm_L"]) and eq(R, g["sgm_R"])
    # quirk 1: equals the chained passes, NOT the independent-direction average
    x = g["cbca1_L"].copy()
    for r in DIRS:
        oracle.semi_global_matching(g["left_image"], g["right_image"], x, r, 2.3 if r[0] == 0 else 2.3 / 1.5,
                                    55.9, 4, 8, 0.08, "L")
    assert eq(x, g["sgm_L"])


def test_wta_and_refinement_bit_exact(oracle, pipeline_golden):
    g = pipeline_golden
    D = int(g["ndisp"])
    dl, dr = oracle.disparity_prediction(g["cbca2_L"], g["cbca2_R"])
    assert eq(dl, g["wta_L"]) and eq(dr, g["wta_R"]) and dl.dtype == np.float32
    d = oracle.interpolation(dl, dr, D)
    assert eq(d, g["interp"])
    d = oracle.subpixel_enhance(d, g["cbca2_L"])
    assert eq(d, g["subpixel"])
    d = oracle.median_filter(d, 5, 5)
    assert eq(d, g["median"])
    d = oracle.bilateral_filter(g["left_image"], d, 5, 5, 0, 6, 2)
    assert eq(d, g["bilateral"])


def test_whole_pipeline_from_features(oracle, pipeline_golden):
    g = pipeline_golden
    d = oracle.match_from_features(g["left_image"], g["right_image"], g["fl"], g["fr"], int(g["ndisp"]))
    assert eq(d, g["bilateral"])


def test_integer_cost_cases(oracle, integer_golden):
    g = integer_golden
    dl, dr = oracle.disparity_prediction(g["L"], g["R"])
    assert eq(dl, g["wta_L"]) and eq(dr, g["wta_R"])
    # first-minimum rule on ties == np.argmin
    assert eq(dl, np.argmin(g["L"], axis=0).astype(np.float32))
    for r in DIRS:
        for ch in "LR":
            x = g["R"].copy()
            oracle.semi_global_matching(g["left_image"], g["right_image"], x, r, 2.0, 56.0, 4, 8, 0.08, ch)
            assert eq(x, g["sgm_int_%s_%d_%d" % (ch, r[0], r[1])])
    D = g["L"].shape[0]
    assert eq(oracle.interpolation(g["rand_dl"], g["rand_dr"], D), g["rand_interp"])
    assert eq(oracle.subpixel_enhance(g["half_disp"], g["sub_vol"]), g["half_subpixel"])
    assert eq(oracle.subpixel_enhance(g["rand_dl"], g["L"]), g["int_subpixel"])   # inf / NaN cells included
    assert np.isnan(g["int_subpixel"]).any() or np.isinf(g["int_subpixel"]).any()
    assert eq(oracle.median_filter(g["half_subpixel"], 5, 5), g["half_median"])
    assert eq(oracle.bilateral_filter(g["left_image"], g["half_median"], 5, 5, 0, 6, 2), g["half_bilateral"])


def test_flat_image_regions_and_aggregation(oracle, flat_golden):
    """Worst case of pf:585-599: arms at the distance limit, regions of up to 27 x 27 = 729 pixels."""
    g = flat_golden
    region, num = oracle.compute_cross_region(g["left_image"], 0.02, 14)
    assert num.max() == 729 and num.mean() > 400
    assert eq(num, g["region_num_left"])
    assert eq(region[13:15], g["region_left_rows13_15"].astype(np.int32))
    _, numr = oracle.compute_cross_region(g["right_image"], 0.02, 14)
    assert eq(numr, g["region_num_right"])
    for iters in (1, 2, 5):
        L, R = oracle.cost_volume_aggregation(g["left_image"], g["right_image"], g["cv_L"], g["cv_R"], 0.02, 14, iters)
        assert eq(L, g["cbca%d_L" % iters]) and eq(R, g["cbca%d_R" % iters]), iters


def test_feature_net_on_the_shipped_checkpoint(oracle, features_golden, checkpoint_golden):
    """The C restatement of model.py:40-64 with the reference's real weights against the float64 torch restatement."""
    g = features_golden
    ws, bs = checkpoint_golden
    assert eq(ws[0], g["conv1_weights"]) and eq(bs[0], g["conv1_biases"]) and eq(bs[4], g["conv5_biases"])
    np.testing.assert_allclose([float(w.astype(np.float64).sum()) for w in ws], g["weight_sums"], rtol=0, atol=1e-9)
    f = oracle.net_forward(g["image"], ws, bs)
    np.testing.assert_allclose(f, g["features"], atol=2e-6, rtol=0)


def test_feature_net_matches_torch_restatement(oracle, features_golden):
    g = features_golden
    ws, bs = oracle.glorot_uniform_weights(seed=int(g["glorot_seed"]))
    f = oracle.net_forward(g["image"], ws, bs)
    assert f.shape == g["features_glorot"].shape
    np.testing.assert_allclose(f, g["features_glorot"], atol=2e-6, rtol=0)
    np.testing.assert_allclose(np.linalg.norm(f, axis=-1), 1.0, atol=1e-5)


def test_edge_cases(oracle):
    rng = np.random.default_rng(0)
    # smallest legal problem: ndisp = 2 (SGM needs >= 2), W = ndisp + 2
    fl = rng.standard_normal((3, 4, 64)).astype(np.float32)
    fr = rng.standard_normal((3, 4, 64)).astype(np.float32)
    L, R = oracle.compute_cost_volume(fl, fr, 2)
    assert L.shape == (2, 3, 4) and np.isfinite(L).all() and np.isfinite(R).all()
    img = np.zeros((3, 4, 1), np.float32)          # constant image: arms limited only by borders
    arms, cnt = oracle.cross_arms(img, 0.02, 14)
    assert cnt.min() == 12 and cnt.max() == 12
    out = oracle.cbca_one(img, L, 0.02, 14, 1)
    np.testing.assert_allclose(out, np.broadcast_to(L.mean(axis=(1, 2), keepdims=True), L.shape), rtol=1e-5,
                               atol=1e-6)
    # 0 iterations is the identity
    assert np.array_equal(oracle.cbca_one(img, L, 0.02, 14, 0), L)
    # 1x1 disparity maps through the refinement chain
    one = np.array([[1.0]], np.float32)
    assert oracle.median_filter(one, 5, 5)[0, 0] == 1.0
    assert oracle.bilateral_filter(np.zeros((1, 1, 1), np.float32), one, 5, 5, 0, 6, 2)[0, 0] == 1.0
