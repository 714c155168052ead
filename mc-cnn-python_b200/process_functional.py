"""Drop-in for the reference's ``src/process_functional.py`` ("pf"): same function names,
positional argument order, return arity, shapes and dtypes, so that ``match.py``'s
``from process_functional import *`` (match.py:13) and its ten calls (match.py:132-175) work
unchanged.  Every body is a call through the C ABI (include/mccnn_b200.h) into hand-written
sm_100a CUDA kernels; there is no CPU path.

Arrays may be NumPy (copied to the device and back: the reference's exact interface) or CUDA
torch tensors (stay on the device; results are returned as tensors).  Cost volumes have the
reference's LOGICAL shape [ndisp, H, W]; on the device they are stored disparity-fastest
("HWD", see the header), so tensor results are permuted views.  A [D,H,W]-contiguous tensor or
array handed in is converted on the fly.
"""
import ctypes

import numpy as np

try:
    from . import _ffi
    from . import checkpoint as _checkpoint
except ImportError:      # imported as a top-level module with this directory on sys.path (like the reference's src/)
    import _ffi
    import checkpoint as _checkpoint

__all__ = ["compute_features", "compute_cost_volume", "cost_volume_aggregation", "SGM_average",
           "disparity_prediction", "interpolation", "subpixel_enhance", "median_filter", "bilateral_filter",
           "semi_global_matching", "compute_cross_region"]


# Summation order of cross-based aggregation (mccnn_cbca `mode`):
#   0 = "separable" (default): row sums re-used down each column, <= 54 additions per cell; differs
#       from the reference only by float32 re-association (~1e-7 relative); the rounds of a call are
#       chained (column pass of round k + row pass of round k+1 in one kernel, 8 B/cell/round);
#   1 = "exact": the reference's flat running sum over the whole region (pf:157-161), bit-identical
#       to the reference, <= 729 additions per cell;
#   2 = "separable, two passes per round": the sums of mode 0, identical bits, 16 B/cell/round -- the
#       better choice for piece-wise constant images (every vertical arm at the limit), and the cross-check.
CBCA_SEPARABLE, CBCA_EXACT, CBCA_SEPARABLE_TWO_PASS = 0, 1, 2
# Host-side default: pick 0 or 2 per image from the arms (the two are bit-identical, only their speed differs).  The
# chained kernel gathers the rows beyond +-1 of a vertical arm from global memory behind a CTA barrier and recomputes
# the horizontal halo of its segments, both of which grow with the arms: measured at 1024x1024x192, ms per round,
# natural image (mean up + down = 0.96) 0.29 chained / 0.58 two passes, piece-wise constant image (22.1) 2.20 / 1.85;
# linear in the mean, the two meet at 10.
CBCA_AUTO = -1
CBCA_AUTO_MEAN_VERTICAL_ARMS = 10.0
CBCA_MODE = CBCA_AUTO


def cbca_auto_mode(arms):
    """CBCA_SEPARABLE (chained rounds) for natural images, CBCA_SEPARABLE_TWO_PASS when the vertical arms are long.
    Reads one number back from the device (a host synchronisation of a few tens of microseconds per image)."""
    H, W = int(arms.shape[0]), int(arms.shape[1])
    total = _torch().empty(1, dtype=_torch().int64, device=arms.device)
    _ffi.call("mccnn_arms_vertical_sum", _ffi.ptr(arms), H, W, _ffi.ptr(total), _ffi.stream_ptr())
    mean_vertical = float(total.item()) / float(H * W)
    return CBCA_SEPARABLE if mean_vertical < CBCA_AUTO_MEAN_VERTICAL_ARMS else CBCA_SEPARABLE_TWO_PASS


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise _ffi.MccnnError("a CUDA device is required (there is no CPU fallback)")
    return torch


def _dev():
    return _torch().device("cuda", _torch().cuda.current_device())


def _is_tensor(x):
    return type(x).__module__.startswith("torch")


def _to_dev(x, dtype=None):
    """NumPy / tensor -> contiguous float32 CUDA tensor."""
    torch = _torch()
    dtype = dtype or torch.float32
    if _is_tensor(x):
        return x.to(device=_dev(), dtype=dtype).contiguous()
    return torch.from_numpy(np.ascontiguousarray(x)).to(device=_dev(), dtype=dtype).contiguous()


def _image2d(image):
    """[H,W,1] or [H,W] -> contiguous [H,W] float32 CUDA tensor."""
    if image.ndim == 3:
        assert image.shape[2] == 1, "expected a single-channel image [H,W,1]"
        image = image[:, :, 0]
    assert image.ndim == 2
    return _to_dev(image)


def _ret(t, like):
    """Return `t` as the kind of array `like` was."""
    if _is_tensor(like):
        return t
    return t.cpu().numpy()


# ------------------------------------------------------------------------------------------ volumes
def _hwd_view(hwd, D):
    """[H,W,Dp] storage -> logical [D,H,W] view (no copy)."""
    return hwd[:, :, :D].permute(2, 0, 1)


def _is_hwd_view(vol):
    if not _is_tensor(vol) or not vol.is_cuda or vol.ndim != 3 or vol.dtype != _torch().float32:
        return False
    D, H, W = vol.shape
    Dp = _ffi.dpitch(D)
    return vol.stride() == (1, W * Dp, Dp) and vol.data_ptr() % 16 == 0


def _as_hwd(vol):
    """Logical [D,H,W] volume (NumPy, DHW tensor, or one of our HWD views) -> (hwd [H,W,Dp], D, H, W)."""
    torch = _torch()
    assert vol.ndim == 3, "cost volume must be [ndisp, H, W]"
    D, H, W = (int(s) for s in vol.shape)
    Dp = _ffi.dpitch(D)
    if _is_hwd_view(vol):
        return torch.as_strided(vol, (H, W, Dp), (W * Dp, Dp, 1)), D, H, W
    dhw = _to_dev(vol)
    hwd = torch.empty((H, W, Dp), dtype=torch.float32, device=dhw.device)
    _ffi.call("mccnn_dhw_to_hwd", _ffi.ptr(dhw), _ffi.ptr(hwd), D, H, W, _ffi.stream_ptr())
    return hwd, D, H, W


def _hwd_to_dhw(hwd, D):
    torch = _torch()
    H, W, _ = hwd.shape
    dhw = torch.empty((D, H, W), dtype=torch.float32, device=hwd.device)
    _ffi.call("mccnn_hwd_to_dhw", _ffi.ptr(hwd), _ffi.ptr(dhw), D, H, W, _ffi.stream_ptr())
    return dhw


def _ret_volume(hwd, D, like):
    if _is_tensor(like):
        return _hwd_view(hwd, D)
    return _hwd_to_dhw(hwd, D).cpu().numpy()


def _empty_hwd(H, W, D):
    torch = _torch()
    return torch.empty((H, W, _ffi.dpitch(D)), dtype=torch.float32, device=_dev())


# ------------------------------------------------------------------------------------------ a1/a2
_weight_cache = {}


def glorot_uniform_weights(seed=0, num_layers=5, feature_maps=64, kernel=3, in_channels=1):
    """The reference's random initialisation (tf.get_variable default = glorot-uniform for weights
    AND biases, model.py:100-101), seeded; what `checkpoint=None` uses."""
    rng = np.random.default_rng(seed)
    ws, bs = [], []
    ic = in_channels
    for _ in range(num_layers):
        fan_in, fan_out = kernel * kernel * ic, kernel * kernel * feature_maps
        lim = np.sqrt(6.0 / (fan_in + fan_out))
        ws.append(rng.uniform(-lim, lim, (kernel, kernel, ic, feature_maps)).astype(np.float32))
        blim = np.sqrt(6.0 / (feature_maps + feature_maps))
        bs.append(rng.uniform(-blim, blim, (feature_maps,)).astype(np.float32))
        ic = feature_maps
    return ws, bs


class DeviceWeights(object):
    """conv weights/biases resident on the device + the host pointer tables the C ABI takes."""

    def __init__(self, weights, biases):
        assert len(weights) == len(biases) and len(weights) >= 1
        F = int(weights[0].shape[-1])
        assert F == 64, "the CUDA path implements 64 feature maps (model.py:38)"
        ic = 1
        for w, b in zip(weights, biases):
            assert tuple(w.shape) == (3, 3, ic, F), "weights must be HWIO [3,3,%d,%d], got %r" % (ic, F, tuple(w.shape))
            assert tuple(b.shape) == (F,)
            ic = F
        self.num_layers = len(weights)
        self.w = [_to_dev(w) for w in weights]
        self.b = [_to_dev(b) for b in biases]
        n = self.num_layers
        self.w_table = (ctypes.c_void_p * n)(*[w.data_ptr() for w in self.w])
        self.b_table = (ctypes.c_void_p * n)(*[b.data_ptr() for b in self.b])
        # the constant part of the network, once: hi/lo tf32 split of the weights of layers 2..n (mccnn_features_prepare)
        nb = int(_ffi.lib().mccnn_features_weights_bytes(n))
        self.prepared = _torch().empty((max(nb, 32) + 3) // 4, dtype=_torch().float32, device=self.w[0].device)
        _ffi.call("mccnn_features_prepare", n, self.w_table, _ffi.ptr(self.prepared), _ffi.stream_ptr())
        _torch().cuda.current_stream().synchronize()        # (used from whichever stream calls the network later)


def resolve_weights(checkpoint, num_layers=5):
    """checkpoint: TF bundle prefix (as `--resume`, pf:32/:43) | (weights, biases) | DeviceWeights | None."""
    if isinstance(checkpoint, DeviceWeights):
        return checkpoint
    if checkpoint is None:
        key = ("glorot", num_layers, _torch().cuda.current_device())
        if key not in _weight_cache:
            _weight_cache[key] = DeviceWeights(*glorot_uniform_weights(0, num_layers))
        return _weight_cache[key]
    if isinstance(checkpoint, (tuple, list)):
        return DeviceWeights(checkpoint[0], checkpoint[1])
    key = (str(checkpoint), _torch().cuda.current_device())
    if key not in _weight_cache:
        ws, bs = _checkpoint.load_mccnn_weights(str(checkpoint))
        _weight_cache[key] = DeviceWeights(ws, bs)
    return _weight_cache[key]


def net_forward(image2d, dw, pad, out=None, scratch=None):
    """One image [H,W] (device) through the network; implicit zero padding `pad`."""
    torch = _torch()
    H, W = (int(s) for s in image2d.shape)
    n = dw.num_layers
    OH, OW = H + 2 * pad - 2 * n, W + 2 * pad - 2 * n
    assert OH >= 1 and OW >= 1, "image too small for %d VALID 3x3 layers" % n
    if out is None:
        out = torch.empty((OH, OW, 64), dtype=torch.float32, device=image2d.device)
    nbytes = int(_ffi.lib().mccnn_features_scratch_bytes(H, W, pad, n))
    if scratch is None or scratch.numel() * scratch.element_size() < nbytes:
        scratch = torch.empty((max(nbytes, 4) + 3) // 4, dtype=torch.float32, device=image2d.device)
    _ffi.call("mccnn_features_prepared", _ffi.ptr(image2d), H, W, pad, n, dw.w_table, dw.b_table, _ffi.ptr(dw.prepared),
              _ffi.ptr(out), _ffi.ptr(scratch), _ffi.stream_ptr())
    return out


def compute_features(left_image, right_image, patch_height, patch_width, checkpoint):
    """pf:15.  Returns (featuresl, featuresr), each [H, W, 64] float32, L2-normalised."""
    ph, pw = int(patch_height), int(patch_width)
    assert ph == pw and ph % 2 == 1, "square odd patch expected (match.py:132 passes 11, 11)"
    pad = (ph - 1) // 2                                           # pf:22-23
    dw = resolve_weights(checkpoint, num_layers=pad)
    assert dw.num_layers == pad, "patch size %d needs %d conv layers, checkpoint has %d" % (ph, pad, dw.num_layers)
    li, ri = _image2d(left_image), _image2d(right_image)
    assert li.shape == ri.shape
    fl = net_forward(li, dw, pad)
    fr = net_forward(ri, dw, pad)
    return _ret(fl, left_image), _ret(fr, right_image)


# ------------------------------------------------------------------------------------------ a3
def compute_cost_volume(featuresl, featuresr, ndisp):
    """pf:78.  Returns (left_cost_volume, right_cost_volume), each [ndisp, H, W] float32."""
    ndisp = int(ndisp)
    fl, fr = _to_dev(featuresl), _to_dev(featuresr)
    assert fl.ndim == 3 and fl.shape == fr.shape
    H, W, C = (int(s) for s in fl.shape)
    L, R = _empty_hwd(H, W, ndisp), _empty_hwd(H, W, ndisp)
    _ffi.call("mccnn_cost_volume", _ffi.ptr(fl), _ffi.ptr(fr), _ffi.ptr(L), _ffi.ptr(R), H, W, C, ndisp,
              _ffi.stream_ptr())
    return _ret_volume(L, ndisp, featuresl), _ret_volume(R, ndisp, featuresr)


# ------------------------------------------------------------------------------------------ a4/a5
def cross_arms(image, intensity_threshold, distance_threshold):
    """Arm-length form of pf:571: (arms u8 [H,W,4] = up, down, left, right ; count i32 [H,W]) on the device."""
    torch = _torch()
    img = _image2d(image)
    H, W = (int(s) for s in img.shape)
    dist = int(distance_threshold)
    assert dist == distance_threshold, "distance_threshold must be integer-valued (SURVEY.md section 5 config trap)"
    arms = torch.empty((H, W, 4), dtype=torch.uint8, device=img.device)
    count = torch.empty((H, W), dtype=torch.int32, device=img.device)
    _ffi.call("mccnn_cross_arms", _ffi.ptr(img), _ffi.ptr(arms), _ffi.ptr(count), H, W,
              ctypes.c_float(np.float32(intensity_threshold)), dist, _ffi.stream_ptr())
    return arms, count


def compute_cross_region(image, intensity_threshold, distance_threshold):
    """pf:571 compatibility view: (union_region [H,W,(2*dist)^2,2] int32 padded with -1, union_region_num [H,W])."""
    torch = _torch()
    arms, count = cross_arms(image, intensity_threshold, distance_threshold)
    H, W = (int(s) for s in count.shape)
    dist = int(distance_threshold)
    region = torch.empty((H, W, (2 * dist) ** 2, 2), dtype=torch.int32, device=arms.device)
    _ffi.call("mccnn_cross_region_list", _ffi.ptr(arms), _ffi.ptr(region), H, W, dist, _ffi.stream_ptr())
    return _ret(region, image), _ret(count, image)


def _cbca_one(hwd, D, arms, count, iters, dist, out=None, scratch=None, mode=None):
    H, W, _ = hwd.shape
    if out is None:
        out = _empty_hwd(H, W, D)
    if mode is None:
        mode = CBCA_MODE
    if mode == CBCA_AUTO:
        mode = cbca_auto_mode(arms)
    if scratch is None and iters >= 1:
        scratch = _empty_hwd(H, W, D)
    _ffi.call("mccnn_cbca", _ffi.ptr(hwd), _ffi.ptr(out), _ffi.ptr(scratch), _ffi.ptr(arms), _ffi.ptr(count),
              D, int(H), int(W), int(iters), int(dist), int(mode), _ffi.stream_ptr())
    return out


def cost_volume_aggregation(left_image, right_image, left_cost_volume, right_cost_volume,
                            intensity_threshold, distance_threshold, max_average_time):
    """pf:117.  Fresh result volumes; the inputs are left untouched."""
    iters = int(max_average_time)
    assert iters == max_average_time and iters >= 0
    outs = []
    scratch = None
    for image, vol in ((left_image, left_cost_volume), (right_image, right_cost_volume)):
        arms, count = cross_arms(image, intensity_threshold, distance_threshold)      # pf:120-121
        hwd, D, H, W = _as_hwd(vol)
        assert tuple(count.shape) == (H, W), "image and cost volume shapes differ"
        if iters >= 1 and (scratch is None or scratch.shape != hwd.shape):
            scratch = _empty_hwd(H, W, D)
        out = _cbca_one(hwd, D, arms, count, iters, int(distance_threshold), scratch=scratch)
        outs.append(_ret_volume(out, D, vol))
    return outs[0], outs[1]


# ------------------------------------------------------------------------------------------ a6/a7
def _sgm_scratch(H, W, D):
    torch = _torch()
    n = int(_ffi.lib().mccnn_sgm_scratch_bytes(H, W, D))
    return torch.empty((n + 3) // 4, dtype=torch.int32, device=_dev())


def semi_global_matching(left_image, right_image, cost_volume, r, sgm_P1, sgm_P2, sgm_Q1, sgm_Q2, sgm_D, choice):
    """pf:476.  Updates `cost_volume` IN PLACE and returns the same object (the reference's aliasing, pf:544)."""
    assert choice == "R" or choice == "L"                        # pf:479
    assert r[0] * r[1] == 0                                      # pf:484
    li, ri = _image2d(left_image), _image2d(right_image)
    hwd, D, H, W = _as_hwd(cost_volume)
    assert tuple(li.shape) == (H, W) and tuple(ri.shape) == (H, W)
    flags = _sgm_scratch(H, W, D)
    _ffi.call("mccnn_sgm_pass", _ffi.ptr(hwd), _ffi.ptr(li), _ffi.ptr(ri), _ffi.ptr(flags), D, H, W,
              int(r[0]), int(r[1]), float(sgm_P1), float(sgm_P2), float(sgm_Q1), float(sgm_Q2), float(sgm_D),
              1 if choice == "L" else 0, _ffi.stream_ptr())
    if _is_hwd_view(cost_volume):
        return cost_volume                                        # updated in place, zero copy
    dhw = _hwd_to_dhw(hwd, D)
    if _is_tensor(cost_volume):
        cost_volume.copy_(dhw)
    else:
        np.copyto(cost_volume, dhw.cpu().numpy())
    return cost_volume


def SGM_average(left_cost_volume, right_cost_volume, left_image, right_image,
                sgm_P1, sgm_P2, sgm_Q1, sgm_Q2, sgm_D, sgm_V):
    """pf:187.  Four chained in-place passes per volume, (0,1) (0,-1) (-1,0) (1,0) (pf:194-208); the
    reference's closing (X+X+X+X)/4. is the identity.  Like the reference (pf:195-232) the
    volumes handed in are MUTATED -- they hold the chained passes' result afterwards -- and the returned volumes are
    equal to them: NumPy arrays are written back and fresh arrays returned; HWD-view tensors are updated in place
    (the returned views alias them, zero copy); DHW tensors are written back."""
    li, ri = _image2d(left_image), _image2d(right_image)
    hl, D, H, W = _as_hwd(left_cost_volume)
    hr, D2, H2, W2 = _as_hwd(right_cost_volume)
    assert (D, H, W) == (D2, H2, W2) and tuple(li.shape) == (H, W) and tuple(ri.shape) == (H, W)
    flags = _sgm_scratch(H, W, D)
    _ffi.call("mccnn_sgm_average_pair", _ffi.ptr(hl), _ffi.ptr(hr), _ffi.ptr(li), _ffi.ptr(ri),
              _ffi.ptr(flags), D, H, W, float(sgm_P1), float(sgm_P2), float(sgm_Q1), float(sgm_Q2),
              float(sgm_D), float(sgm_V), _ffi.stream_ptr())
    outs = _ret_volume(hl, D, left_cost_volume), _ret_volume(hr, D, right_cost_volume)
    for given, out in zip((left_cost_volume, right_cost_volume), outs):
        if isinstance(given, np.ndarray):
            np.copyto(given, out)                                 # the reference's in-place passes (pf:544)
        elif _is_tensor(given) and not _is_hwd_view(given):
            given.copy_(out)
    return outs


# ------------------------------------------------------------------------------------------ a8
def _wta(hwd, D, out=None):
    torch = _torch()
    H, W, _ = hwd.shape
    if out is None:
        out = torch.empty((H, W), dtype=torch.float32, device=hwd.device)
    _ffi.call("mccnn_wta", _ffi.ptr(hwd), _ffi.ptr(out), D, int(H), int(W), _ffi.stream_ptr())
    return out


def disparity_prediction(left_cost_volume, right_cost_volume):
    """pf:239.  First-minimum argmin over d; float32 maps [H,W]."""
    hl, D, _, _ = _as_hwd(left_cost_volume)
    hr, D2, _, _ = _as_hwd(right_cost_volume)
    return _ret(_wta(hl, D), left_cost_volume), _ret(_wta(hr, D2), right_cost_volume)


# ------------------------------------------------------------------------------------------ a9-a12
def interpolation(left_disparity_map, right_disparity_map, ndisp, return_labels=False):
    """pf:279."""
    torch = _torch()
    dl, dr = _to_dev(left_disparity_map), _to_dev(right_disparity_map)
    assert dl.ndim == 2 and dl.shape == dr.shape
    H, W = (int(s) for s in dl.shape)
    out = torch.empty_like(dl)
    labels = torch.empty((H, W), dtype=torch.int32, device=dl.device)
    _ffi.call("mccnn_lr_interp", _ffi.ptr(dl), _ffi.ptr(dr), _ffi.ptr(out), _ffi.ptr(labels), H, W, int(ndisp),
              _ffi.stream_ptr())
    if return_labels:
        return _ret(out, left_disparity_map), _ret(labels, left_disparity_map)
    return _ret(out, left_disparity_map)


def subpixel_enhance(left_disparity_map, left_cost_volume):
    """pf:381."""
    torch = _torch()
    d = _to_dev(left_disparity_map)
    hwd, D, H, W = _as_hwd(left_cost_volume)
    assert tuple(d.shape) == (H, W)
    out = torch.empty_like(d)
    _ffi.call("mccnn_subpixel", _ffi.ptr(d), _ffi.ptr(hwd), _ffi.ptr(out), D, H, W, _ffi.stream_ptr())
    return _ret(out, left_disparity_map)


def median_filter(left_disparity_map, filter_height, filter_width):
    """pf:403."""
    torch = _torch()
    d = _to_dev(left_disparity_map)
    H, W = (int(s) for s in d.shape)
    out = torch.empty_like(d)
    _ffi.call("mccnn_median", _ffi.ptr(d), _ffi.ptr(out), H, W, int(filter_height), int(filter_width), _ffi.stream_ptr())
    return _ret(out, left_disparity_map)


def bilateral_table(filter_height, filter_width, mean, std_dev):
    """The float32 weight table of pf:428-436: util.normal (util.py:45-48) evaluated in float64 at the
    Euclidean distance from the window centre, stored as float32."""
    constant1 = 1. / (np.sqrt(2 * np.pi) * std_dev)
    constant2 = -1. / (2 * std_dev * std_dev)
    ch, cw = (filter_height - 1) // 2, (filter_width - 1) // 2
    t = np.zeros([filter_height, filter_width], dtype=np.float32)
    for h in range(filter_height):
        for w in range(filter_width):
            x = np.sqrt((h - ch) ** 2 + (w - cw) ** 2)
            t[h, w] = constant1 * np.exp(constant2 * ((x - mean) ** 2))
    return t


_table_cache = {}


def bilateral_filter(left_image, left_disparity_map, filter_height, filter_width, mean, std_dev, blur_threshold):
    """pf:424."""
    torch = _torch()
    fh, fw = int(filter_height), int(filter_width)
    img = _image2d(left_image)
    d = _to_dev(left_disparity_map)
    H, W = (int(s) for s in d.shape)
    assert tuple(img.shape) == (H, W)
    key = (fh, fw, float(mean), float(std_dev), torch.cuda.current_device())
    if key not in _table_cache:
        _table_cache[key] = _to_dev(bilateral_table(fh, fw, mean, std_dev))
    out = torch.empty_like(d)
    _ffi.call("mccnn_bilateral", _ffi.ptr(img), _ffi.ptr(d), _ffi.ptr(out), _ffi.ptr(_table_cache[key]), H, W, fh, fw,
              ctypes.c_float(np.float32(blur_threshold)), _ffi.stream_ptr())
    return _ret(out, left_disparity_map)
