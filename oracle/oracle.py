"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

ctypes front-end of ``oracle/mccnn_oracle.c`` (the CPU restatement of the reference's
hot path) exposing the reference's own function names and signatures
(/root/reference/src/process_functional.py), NumPy in / NumPy out, so parity tests read
like calls into the reference.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this module.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmccnn_oracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_u8p = ctypes.POINTER(ctypes.c_uint8)
_i32p = ctypes.POINTER(ctypes.c_int32)


def build(force=False):
    """Compile the C oracle (gcc, a few seconds).  Building the checker is not using it."""
    src = os.path.join(_HERE, "mccnn_oracle.c")
    if force or not os.path.isfile(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libmccnn_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.mccnn_oracle_num_threads.restype = ctypes.c_int
    return _lib


def num_threads():
    return int(lib().mccnn_oracle_num_threads())


def set_threads(n):
    lib().mccnn_oracle_set_threads(ctypes.c_int(int(n)))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t=_f32p):
    return a.ctypes.data_as(t)


def _img2d(image):
    image = np.asarray(image)
    if image.ndim == 3:
        assert image.shape[2] == 1
        image = image[:, :, 0]
    return _f32(image)


# ---------------------------------------------------------------------------- features
def glorot_uniform_weights(seed=0, num_layers=5, feature_maps=64, kernel=3, in_channels=1):
    """Random-init weights as the reference's tf.get_variable default produces them
    (glorot-uniform for weights AND biases, model.py:100-101; limits in SURVEY.md section 8d)."""
    rng = np.random.default_rng(seed)
    ws, bs = [], []
    ic = in_channels
    for _ in range(num_layers):
        fan_in, fan_out = kernel * kernel * ic, kernel * kernel * feature_maps
        lim = np.sqrt(6.0 / (fan_in + fan_out))
        ws.append(rng.uniform(-lim, lim, (kernel, kernel, ic, feature_maps)).astype(np.float32))
        blim = np.sqrt(6.0 / (feature_maps + feature_maps))   # 1-D shape [F]: fan_in = fan_out = F
        bs.append(rng.uniform(-blim, blim, (feature_maps,)).astype(np.float32))
        ic = feature_maps
    return ws, bs


def net_forward(image, weights, biases):
    """NET(x).features for one image [H,W] or [H,W,1] -> [H,W,F] (model.py:40-64, pf:15-73)."""
    img = _img2d(image)
    H, W = img.shape
    nl = len(weights)
    F = weights[0].shape[-1]
    ws = [_f32(w) for w in weights]
    bs = [_f32(b) for b in biases]
    assert ws[0].shape == (3, 3, 1, F)
    wp = (_f32p * nl)(*[_p(w) for w in ws])
    bp = (_f32p * nl)(*[_p(b) for b in bs])
    out = np.empty((H, W, F), dtype=np.float32)
    rc = lib().mccnn_oracle_features(_p(img), H, W, nl, F, wp, bp, _p(out))
    assert rc == 0
    return out


def compute_features(left_image, right_image, patch_height, patch_width, weights_biases):
    """pf:15.  ``weights_biases`` = (weights list, biases list) instead of a checkpoint path."""
    weights, biases = weights_biases
    assert (patch_height - 1) // 2 == len(weights) and (patch_width - 1) // 2 == len(weights)
    return net_forward(left_image, weights, biases), net_forward(right_image, weights, biases)


# ---------------------------------------------------------------------------- cost volume
def compute_cost_volume(featuresl, featuresr, ndisp):
    """pf:78."""
    fl, fr = _f32(featuresl), _f32(featuresr)
    H, W, C = fl.shape
    L = np.empty((ndisp, H, W), dtype=np.float32)
    R = np.empty((ndisp, H, W), dtype=np.float32)
    rc = lib().mccnn_oracle_cost_volume(_p(fl), _p(fr), H, W, C, int(ndisp), _p(L), _p(R))
    assert rc == 0, "need W >= ndisp + 2"
    return L, R


# ---------------------------------------------------------------------------- CBCA
def cross_arms(image, intensity_threshold, distance_threshold):
    """Arm-length form of pf:571 compute_cross_region: (arms u8 [H,W,4] = up,down,left,right; count i32 [H,W])."""
    img = _img2d(image)
    H, W = img.shape
    arms = np.empty((H, W, 4), dtype=np.uint8)
    count = np.empty((H, W), dtype=np.int32)
    rc = lib().mccnn_oracle_cross_arms(_p(img), H, W, ctypes.c_float(np.float32(intensity_threshold)),
                                       int(distance_threshold), _p(arms, _u8p), _p(count, _i32p))
    assert rc == 0
    return arms, count


def compute_cross_region(image, intensity_threshold, distance_threshold):
    """pf:571 -- the reference's explicit representation (union_region [H,W,(2*dist)^2,2] int32, num [H,W])."""
    arms, count = cross_arms(image, intensity_threshold, distance_threshold)
    H, W = count.shape
    dist = int(distance_threshold)
    region = np.empty((H, W, (2 * dist) ** 2, 2), dtype=np.int32)
    rc = lib().mccnn_oracle_cross_region_list(_p(arms, _u8p), H, W, dist, _p(region, _i32p))
    assert rc == 0
    return region, count


def cbca_one(image, volume, intensity_threshold, distance_threshold, iters):
    arms, count = cross_arms(image, intensity_threshold, distance_threshold)
    vol = _f32(volume)
    D, H, W = vol.shape
    out = np.empty_like(vol)
    rc = lib().mccnn_oracle_cbca(_p(vol), _p(out), _p(arms, _u8p), _p(count, _i32p), D, H, W, int(iters))
    assert rc == 0
    return out


def cost_volume_aggregation(left_image, right_image, left_cost_volume, right_cost_volume,
                            intensity_threshold, distance_threshold, max_average_time):
    """pf:117."""
    return (cbca_one(left_image, left_cost_volume, intensity_threshold, distance_threshold, max_average_time),
            cbca_one(right_image, right_cost_volume, intensity_threshold, distance_threshold, max_average_time))


# ---------------------------------------------------------------------------- SGM
def semi_global_matching(left_image, right_image, cost_volume, r, sgm_P1, sgm_P2, sgm_Q1, sgm_Q2, sgm_D, choice):
    """pf:476 -- mutates ``cost_volume`` in place and returns it (the reference's aliasing)."""
    assert choice == "R" or choice == "L"
    assert r[0] * r[1] == 0
    assert cost_volume.dtype == np.float32 and cost_volume.flags["C_CONTIGUOUS"]
    li, ri = _img2d(left_image), _img2d(right_image)
    D, H, W = cost_volume.shape
    rc = lib().mccnn_oracle_sgm_pass(_p(cost_volume), _p(li), _p(ri), D, H, W, int(r[0]), int(r[1]),
                                     ctypes.c_double(sgm_P1), ctypes.c_double(sgm_P2), ctypes.c_double(sgm_Q1),
                                     ctypes.c_double(sgm_Q2), ctypes.c_double(sgm_D), 1 if choice == "L" else 0)
    assert rc == 0
    return cost_volume


def SGM_average(left_cost_volume, right_cost_volume, left_image, right_image,
                sgm_P1, sgm_P2, sgm_Q1, sgm_Q2, sgm_D, sgm_V):
    """pf:187 -- four chained in-place passes per volume; mutates the inputs like the reference."""
    li, ri = _img2d(left_image), _img2d(right_image)
    outs = []
    for vol, is_left in ((left_cost_volume, 1), (right_cost_volume, 0)):
        assert vol.dtype == np.float32 and vol.flags["C_CONTIGUOUS"]
        D, H, W = vol.shape
        rc = lib().mccnn_oracle_sgm_average(_p(vol), _p(li), _p(ri), D, H, W,
                                            ctypes.c_double(sgm_P1), ctypes.c_double(sgm_P2),
                                            ctypes.c_double(sgm_Q1), ctypes.c_double(sgm_Q2),
                                            ctypes.c_double(sgm_D), ctypes.c_double(sgm_V), is_left)
        assert rc == 0
        outs.append(vol.copy())   # the reference returns a fresh array ((X+X+X+X)/4.)
    return outs[0], outs[1]


# ---------------------------------------------------------------------------- WTA + refinement
def wta_one(volume):
    vol = _f32(volume)
    D, H, W = vol.shape
    disp = np.empty((H, W), dtype=np.float32)
    rc = lib().mccnn_oracle_wta(_p(vol), D, H, W, _p(disp))
    assert rc == 0
    return disp


def disparity_prediction(left_cost_volume, right_cost_volume):
    """pf:239."""
    return wta_one(left_cost_volume), wta_one(right_cost_volume)


def interpolation(left_disparity_map, right_disparity_map, ndisp, return_labels=False):
    """pf:279."""
    dl, dr = _f32(left_disparity_map), _f32(right_disparity_map)
    H, W = dl.shape
    out = np.empty((H, W), dtype=np.float32)
    labels = np.empty((H, W), dtype=np.int32)
    rc = lib().mccnn_oracle_interpolation(_p(dl), _p(dr), H, W, int(ndisp), _p(out), _p(labels, _i32p))
    assert rc == 0
    return (out, labels) if return_labels else out


def subpixel_enhance(left_disparity_map, left_cost_volume):
    """pf:381."""
    d, vol = _f32(left_disparity_map), _f32(left_cost_volume)
    D, H, W = vol.shape
    out = np.empty((H, W), dtype=np.float32)
    rc = lib().mccnn_oracle_subpixel(_p(d), _p(vol), D, H, W, _p(out))
    assert rc == 0
    return out


def median_filter(left_disparity_map, filter_height, filter_width):
    """pf:403."""
    d = _f32(left_disparity_map)
    H, W = d.shape
    out = np.empty((H, W), dtype=np.float32)
    rc = lib().mccnn_oracle_median(_p(d), H, W, int(filter_height), int(filter_width), _p(out))
    assert rc == 0
    return out


def bilateral_table(filter_height, filter_width, mean, std_dev):
    """The float32 weight table of pf:428-436 (util.normal evaluated in float64, stored float32)."""
    constant1 = 1. / (np.sqrt(2 * np.pi) * std_dev)
    constant2 = -1. / (2 * std_dev * std_dev)
    ch, cw = (filter_height - 1) // 2, (filter_width - 1) // 2
    t = np.zeros([filter_height, filter_width], dtype=np.float32)
    for h in range(filter_height):
        for w in range(filter_width):
            x = np.sqrt((h - ch) ** 2 + (w - cw) ** 2)
            t[h, w] = constant1 * np.exp(constant2 * ((x - mean) ** 2))
    return t


def bilateral_filter(left_image, left_disparity_map, filter_height, filter_width, mean, std_dev, blur_threshold):
    """pf:424."""
    img, d = _img2d(left_image), _f32(left_disparity_map)
    H, W = d.shape
    table = bilateral_table(int(filter_height), int(filter_width), mean, std_dev)
    out = np.empty((H, W), dtype=np.float32)
    rc = lib().mccnn_oracle_bilateral(_p(img), _p(d), H, W, int(filter_height), int(filter_width), _p(table),
                                      ctypes.c_float(np.float32(blur_threshold)), _p(out))
    assert rc == 0
    return out


# ---------------------------------------------------------------------------- whole pipeline
DEFAULTS = dict(cbca_intensity=0.02, cbca_distance=14, cbca_num_iterations1=2, cbca_num_iterations2=16,
                sgm_P1=2.3, sgm_P2=55.9, sgm_Q1=4, sgm_Q2=8, sgm_D=0.08, sgm_V=1.5,
                blur_sigma=6, blur_threshold=2)


def match_from_features(left_image, right_image, fl, fr, ndisp, return_stages=False, **hp):
    """match.py:137-175 after compute_features, default hyper-parameters of match.py:32-43."""
    p = dict(DEFAULTS)
    p.update(hp)
    st = {}
    L, R = compute_cost_volume(fl, fr, ndisp)
    st["cost_volume"] = (L.copy(), R.copy())
    L, R = cost_volume_aggregation(left_image, right_image, L, R, p["cbca_intensity"], p["cbca_distance"],
                                   p["cbca_num_iterations1"])
    st["cbca1"] = (L.copy(), R.copy())
    L, R = SGM_average(L, R, left_image, right_image, p["sgm_P1"], p["sgm_P2"], p["sgm_Q1"], p["sgm_Q2"],
                       p["sgm_D"], p["sgm_V"])
    st["sgm"] = (L.copy(), R.copy())
    L, R = cost_volume_aggregation(left_image, right_image, L, R, p["cbca_intensity"], p["cbca_distance"],
                                   p["cbca_num_iterations2"])
    st["cbca2"] = (L, R)
    dl, dr = disparity_prediction(L, R)
    st["wta"] = (dl, dr)
    d = interpolation(dl, dr, ndisp)
    st["interpolation"] = d
    d = subpixel_enhance(d, L)
    st["subpixel"] = d
    d = median_filter(d, 5, 5)
    st["median"] = d
    d = bilateral_filter(left_image, d, 5, 5, 0, p["blur_sigma"], p["blur_threshold"])
    st["bilateral"] = d
    return (d, st) if return_stages else d
