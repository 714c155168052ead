"""The body of the reference's per-pair loop (match.py:131-175) as one reusable object.

``StereoMatcher`` owns every device buffer a pair of shape (H, W, ndisp) needs (allocated once
with torch), and ``run`` issues the hot path -- features, cost volume, CBCA x iters1, four chained
SGM passes, CBCA x iters2, WTA, LR-check/interpolation, sub-pixel, median, bilateral -- as a fixed
sequence of C-ABI calls on the current CUDA stream: no allocation and no host<->device traffic inside, and one host
synchronisation (a number per image read back after the cross arms to choose between the two bit-identical
aggregation schedules; none when `cbca_mode` is given explicitly).  ``run_host`` is the same with NumPy images in / NumPy disparity out
through pinned staging buffers (what match.py's loop body amounts to end to end).

Hyper-parameter defaults are match.py:32-43.
"""
import ctypes

import numpy as np

try:
    from . import _ffi
    from . import process_functional as _pf
except ImportError:
    import _ffi
    import process_functional as _pf

DEFAULTS = dict(patch_size=11, cbca_intensity=0.02, cbca_distance=14, cbca_num_iterations1=2,
                cbca_num_iterations2=16, sgm_P1=2.3, sgm_P2=55.9, sgm_Q1=4, sgm_Q2=8, sgm_D=0.08, sgm_V=1.5,
                blur_sigma=6, blur_threshold=2)

STAGES = ("features", "cost_volume", "cbca1", "sgm", "cbca2", "wta", "interpolation", "subpixel", "median",
          "bilateral")


class StereoMatcher(object):

    def __init__(self, H, W, ndisp, checkpoint=None, stages=STAGES, cbca_mode=None, fuse_wta=True, **hp):
        torch = _pf._torch()
        self.torch = torch
        self.H, self.W, self.D = int(H), int(W), int(ndisp)
        self.hp = dict(DEFAULTS)
        self.hp.update(hp)
        self.stages = tuple(stages)
        for s in self.stages:
            assert s in STAGES, "unknown stage %r" % s
        H, W, D = self.H, self.W, self.D
        assert D >= 2 and W >= D + 2, "need ndisp >= 2 and W >= ndisp + 2 (pf:94-95, :547-566)"
        assert D <= 512, "ndisp > 512 is not supported by one call (include/mccnn_b200.h); use slab.SlabMatcher"
        dev = _pf._dev()
        self.device = dev
        f32 = torch.float32
        self.pad = (int(self.hp["patch_size"]) - 1) // 2
        self.weights = _pf.resolve_weights(checkpoint, num_layers=self.pad)
        assert self.weights.num_layers == self.pad
        Dp = _ffi.dpitch(D)
        self.Dp = Dp
        e = lambda *shape, dtype=f32: torch.empty(shape, dtype=dtype, device=dev)
        self.img = [e(H, W), e(H, W)]
        self.feat = [e(H, W, 64), e(H, W, 64)]
        nb = int(_ffi.lib().mccnn_features_scratch_bytes(H, W, self.pad, self.pad))
        self.feat_scratch = e((nb + 3) // 4)
        # volumes: A = cost volume / CBCA2 output, B = CBCA1 output / SGM in place, S = CBCA ping-pong scratch
        self.volA = [e(H, W, Dp), e(H, W, Dp)]
        need_b = any(s in self.stages for s in ("cbca1", "sgm", "cbca2"))
        self.volB = [e(H, W, Dp), e(H, W, Dp)] if need_b else [None, None]
        self.volS = e(H, W, Dp) if need_b else None
        self.arms = [e(H, W, 4, dtype=torch.uint8), e(H, W, 4, dtype=torch.uint8)]
        self.count = [e(H, W, dtype=torch.int32), e(H, W, dtype=torch.int32)]
        ns = int(_ffi.lib().mccnn_sgm_scratch_bytes(H, W, D))
        self.sgm_flags = e((ns + 3) // 4, dtype=torch.int32)
        self.disp = [e(H, W), e(H, W)]
        self.tmp = [e(H, W), e(H, W)]
        self.labels = e(H, W, dtype=torch.int32)
        self.table = _pf._to_dev(_pf.bilateral_table(5, 5, 0, self.hp["blur_sigma"]))
        self.host_in = [torch.empty((H, W), dtype=f32, pin_memory=True) for _ in range(2)]
        self.host_out = torch.empty((H, W), dtype=f32, pin_memory=True)
        self.h2d_bytes = 2 * H * W * 4
        self.d2h_bytes = H * W * 4
        self.cbca_mode = _pf.CBCA_MODE if cbca_mode is None else int(cbca_mode)
        self.cbca_modes = [_pf.CBCA_SEPARABLE if self.cbca_mode == _pf.CBCA_AUTO else self.cbca_mode] * 2
        # winner-take-all folded into the closing column pass of the second aggregation (mccnn_cbca_wta: same map, the
        # volumes are not read again); the exact aggregation mode has no such pass and keeps the separate k_wta
        self.fuse_wta = bool(fuse_wta) and "cbca2" in self.stages and "wta" in self.stages and self.cbca_mode != _pf.CBCA_EXACT
        self.wta_keys = e(H, W, dtype=torch.int64) if self.fuse_wta else None
        self.final_volume = None        # HWD left volume the last run's WTA / sub-pixel read
        self.result = None
        self._steps = self._build()

    # ------------------------------------------------------------------------------------------
    def _build(self):
        """List of (stage name, thunk); each thunk makes the C-ABI calls of one stage.  Buffers are
        bound when the thunk is created (factory functions, no late-binding closures)."""
        H, W, D = self.H, self.W, self.D
        hp, p, call, sp = self.hp, _ffi.ptr, _ffi.call, _ffi.stream_ptr
        st = self.stages
        steps = []
        f = ctypes.c_float
        img, feat = self.img, self.feat

        def make_feature(i):
            def feature():
                call("mccnn_features_prepared", p(img[i]), H, W, self.pad, self.pad, self.weights.w_table,
                     self.weights.b_table, p(self.weights.prepared), p(feat[i]), p(self.feat_scratch), sp())
            return feature

        self._feature_fns = [make_feature(0), make_feature(1)]

        def make_features():
            def features():
                for fn in self._feature_fns:
                    fn()
            return features

        def make_cost_volume(vol):
            def cost_volume():
                call("mccnn_cost_volume", p(feat[0]), p(feat[1]), p(vol[0]), p(vol[1]), H, W, 64, D, sp())
            return cost_volume

        def make_arms_of(i):
            def arms_of():
                call("mccnn_cross_arms", p(img[i]), p(self.arms[i]), p(self.count[i]), H, W,
                     f(np.float32(hp["cbca_intensity"])), int(hp["cbca_distance"]), sp())
                # chained rounds or two passes per round: bit-identical, chosen per image from its arms (one number
                # read back from the device: the only host synchronisation of the step)
                self.cbca_modes[i] = _pf.cbca_auto_mode(self.arms[i]) if self.cbca_mode == _pf.CBCA_AUTO else self.cbca_mode
            return arms_of

        self._arms_fns = [make_arms_of(0), make_arms_of(1)]

        def make_arms():
            def arms():
                for fn in self._arms_fns:
                    fn()
            return arms

        def make_cbca(src, dst, iters):
            def cbca():
                for i in range(2):
                    call("mccnn_cbca", p(src[i]), p(dst[i]), p(self.volS), p(self.arms[i]), p(self.count[i]), D, H, W,
                         iters, int(hp["cbca_distance"]), int(self.cbca_modes[i]), sp())
            return cbca

        def make_cbca_wta(src, dst, iters, disp):
            def cbca_wta():
                for i in range(2):
                    call("mccnn_cbca_wta", p(src[i]), p(dst[i]), p(self.volS), p(self.arms[i]), p(self.count[i]), D, H, W,
                         iters, int(hp["cbca_distance"]), int(self.cbca_modes[i]), 1, p(self.wta_keys), p(disp[i]), sp())
            return cbca_wta

        def make_sgm(vol):
            def sgm():
                call("mccnn_sgm_average_pair", p(vol[0]), p(vol[1]), p(img[0]), p(img[1]), p(self.sgm_flags),
                     D, H, W, float(hp["sgm_P1"]), float(hp["sgm_P2"]), float(hp["sgm_Q1"]), float(hp["sgm_Q2"]),
                     float(hp["sgm_D"]), float(hp["sgm_V"]), sp())
            return sgm

        def make_wta(vol, disp):
            def wta():
                for i in range(2):
                    call("mccnn_wta", p(vol[i]), p(disp[i]), D, H, W, sp())
            return wta

        def make_interp(src, right, dst):
            def interp():
                call("mccnn_lr_interp", p(src), p(right), p(dst), p(self.labels), H, W, D, sp())
            return interp

        def make_subpixel(src, vol, dst):
            def subpixel():
                call("mccnn_subpixel", p(src), p(vol), p(dst), D, H, W, sp())
            return subpixel

        def make_median(src, dst):
            def median():
                call("mccnn_median", p(src), p(dst), H, W, 5, 5, sp())
            return median

        def make_bilateral(src, dst):
            def bilateral():
                call("mccnn_bilateral", p(img[0]), p(src), p(dst), p(self.table), H, W, 5, 5,
                     f(np.float32(hp["blur_threshold"])), sp())
            return bilateral

        def other(d):
            return self.tmp[0] if d is not self.tmp[0] else self.tmp[1]

        # the cross arms need the images only: they go first, so that the one host read-back of the step (the choice of the
        # aggregation schedule) happens while the device queue is still empty instead of draining it mid-step
        if "cbca1" in st or "cbca2" in st:
            steps.append(("cross_arms", make_arms()))
        if "features" in st:
            steps.append(("features", make_features()))
        cur = self.volA
        if "cost_volume" in st:
            steps.append(("cost_volume", make_cost_volume(cur)))
        if "cbca1" in st:
            steps.append(("cbca1", make_cbca(self.volA, self.volB, int(hp["cbca_num_iterations1"]))))
            cur = self.volB
        if "sgm" in st:
            steps.append(("sgm", make_sgm(cur)))
        if "cbca2" in st:
            dst = self.volA if cur is self.volB else self.volB
            if self.fuse_wta:
                steps.append(("cbca2", make_cbca_wta(cur, dst, int(hp["cbca_num_iterations2"]), self.disp)))
            else:
                steps.append(("cbca2", make_cbca(cur, dst, int(hp["cbca_num_iterations2"]))))
            cur = dst
        self.final_volume = cur
        d = None
        if "wta" in st:
            if not self.fuse_wta:
                steps.append(("wta", make_wta(cur, self.disp)))
            d = self.disp[0]
        if "interpolation" in st:
            steps.append(("interpolation", make_interp(d, self.disp[1], self.tmp[0])))
            d = self.tmp[0]
        if "subpixel" in st:
            dst = other(d)
            steps.append(("subpixel", make_subpixel(d, cur[0], dst)))
            d = dst
        if "median" in st:
            dst = other(d)
            steps.append(("median", make_median(d, dst)))
            d = dst
        if "bilateral" in st:
            dst = other(d)
            steps.append(("bilateral", make_bilateral(d, dst)))
            d = dst
        self.result = d
        return steps

    # ------------------------------------------------------------------------------------------
    def set_images(self, left_image, right_image):
        """Copy two normalised images ([H,W,1] or [H,W]; NumPy or tensor) into the resident buffers."""
        for i, im in enumerate((left_image, right_image)):
            self.img[i].copy_(_pf._image2d(im))

    def set_features(self, fl, fr):
        self.feat[0].copy_(_pf._to_dev(fl))
        self.feat[1].copy_(_pf._to_dev(fr))

    def run(self):
        """Issue the configured stages on the current stream over the resident inputs; returns the
        device tensor holding the final disparity map (or None if no map-producing stage is enabled)."""
        for _, fn in self._steps:
            fn()
        return self.result

    def run_timed(self):
        """Like run(), with a CUDA-event pair around every stage: returns {stage: milliseconds}."""
        torch = self.torch
        evs = []
        for name, fn in self._steps:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            evs.append((name, a, b))
        torch.cuda.synchronize()
        return {name: a.elapsed_time(b) for name, a, b in evs}

    def run_host(self, left_image, right_image):
        """NumPy images in, NumPy disparity out: H2D from pinned memory, hot path, D2H, one sync."""
        torch = self.torch
        with_features = "features" in self.stages
        with_arms = any(name == "cross_arms" for name, _ in self._steps)
        cur = torch.cuda.current_stream()
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream()
            self._copy_events = [torch.cuda.Event(), torch.cuda.Event()]
        self._copy_stream.wait_stream(cur)          # (earlier work on this stream may still read the images)
        for i, im in enumerate((left_image, right_image)):
            a = np.asarray(im, dtype=np.float32)
            if a.ndim == 3:
                a = a[:, :, 0]
            self.host_in[i].numpy()[...] = a
            with torch.cuda.stream(self._copy_stream):
                self.img[i].copy_(self.host_in[i], non_blocking=True)
                self._copy_events[i].record(self._copy_stream)
            cur.wait_event(self._copy_events[i])
            if with_features:
                self._feature_fns[i]()          # the left image's features run while the right image is staged and copied
            if with_arms:
                # the image's cross arms and the read-back that picks the aggregation schedule ride on the copy stream: the
                # host waits for this image's upload and two small kernels, not for the features queued on the main stream
                with torch.cuda.stream(self._copy_stream):
                    self._arms_fns[i]()
                    self._copy_events[i].record(self._copy_stream)
                cur.wait_event(self._copy_events[i])
        for name, fn in self._steps:
            if name != "features" and name != "cross_arms":
                fn()
        d = self.result
        self.host_out.copy_(d, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self.host_out.numpy().copy()

    def volume(self, which=0):
        """Logical [D,H,W] view of the final left (0) / right (1) cost volume of the last run."""
        return _pf._hwd_view(self.final_volume[which], self.D)


def shard_window(num_pairs, rank, world_size):
    """[start, end) of the pair indices rank `rank` processes: the reference's only parallel mode, disjoint
    `-s/--start`, `-e/--end` windows per process (match.py:26-28, :85-90), made deterministic from (rank, world)."""
    num_pairs, rank, world_size = int(num_pairs), int(rank), int(world_size)
    assert 0 <= rank < world_size and num_pairs >= 0
    base, extra = divmod(num_pairs, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def match_pair(left_image, right_image, ndisp, checkpoint=None, **hp):
    """match.py:131-175 for one pair: normalised images [H,W,1] -> final left disparity map [H,W]."""
    a = left_image
    H, W = int(a.shape[0]), int(a.shape[1])
    m = StereoMatcher(H, W, ndisp, checkpoint=checkpoint, **hp)
    if _pf._is_tensor(left_image):
        m.set_images(left_image, right_image)
        return m.run().clone()
    return m.run_host(left_image, right_image)
