"""ctypes binding of libmccnn_b200.so (include/mccnn_b200.h).

This is the ONLY compute path of the package: there is no CPU or PyTorch fallback.  If the shared
library is missing, or a call fails, an exception is raised.  torch is used for device memory and
streams only; every pointer handed to the library is a raw device address.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmccnn_b200.so")

_c = ctypes
_vp, _i, _d, _f, _sz = _c.c_void_p, _c.c_int, _c.c_double, _c.c_float, _c.c_size_t

# name -> (restype, argtypes); mirrors include/mccnn_b200.h one to one
SIGNATURES = {
    "mccnn_last_error": (_c.c_char_p, []),
    "mccnn_abi_version": (_i, []),
    "mccnn_dpitch": (_i, [_i]),
    "mccnn_launch_count": (_c.c_ulonglong, []),
    "mccnn_dhw_to_hwd": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "mccnn_hwd_to_dhw": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "mccnn_features_scratch_bytes": (_sz, [_i, _i, _i, _i]),
    "mccnn_features": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "mccnn_features_weights_bytes": (_sz, [_i]),
    "mccnn_features_prepare": (_i, [_i, _vp, _vp, _vp]),
    "mccnn_features_prepared": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mccnn_cost_volume": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "mccnn_cross_arms": (_i, [_vp, _vp, _vp, _i, _i, _f, _i, _vp]),
    "mccnn_cross_region_list": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "mccnn_arms_vertical_sum": (_i, [_vp, _i, _i, _vp, _vp]),
    "mccnn_cbca": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "mccnn_cbca_wta": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "mccnn_sgm_scratch_bytes": (_sz, [_i, _i, _i]),
    "mccnn_sgm_pass": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _d, _d, _d, _d, _d, _i, _vp]),
    "mccnn_sgm_average": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _d, _d, _d, _d, _d, _d, _i, _vp]),
    "mccnn_sgm_average_pair": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _d, _d, _d, _d, _d, _d, _vp]),
    "mccnn_wta": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "mccnn_lr_interp": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "mccnn_subpixel": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "mccnn_median": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "mccnn_bilateral": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp]),
    "mccnn_cost_volume_slab": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "mccnn_sgm_passes_slab": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _d, _d, _d, _d, _d, _d, _vp]),
    "mccnn_cbca_to": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _i, _i, _vp]),
    "mccnn_sgm_passes_slab_to": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _d, _d, _d, _d, _d, _d, _i, _vp, _vp, _vp,
                                      _i, _vp]),
    "mccnn_wta_slab": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "mccnn_wta_combine": (_i, [_vp, _vp, _vp, _i, _c.c_longlong, _i, _i, _vp]),
    "mccnn_subpixel_gather": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "mccnn_subpixel_triple": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "mccnn_copy3d": (_i, [_vp, _vp, _c.c_longlong, _c.c_longlong, _i, _c.c_longlong, _c.c_longlong, _c.c_longlong,
                          _c.c_longlong, _vp]),
}

_lib = None


class MccnnError(RuntimeError):
    pass


def lib():
    """Load the library (once).  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise MccnnError("%s is missing: build it with `python mc-cnn-python_b200/build.py` "
                             "(there is no CPU fallback)" % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def last_error():
    return lib().mccnn_last_error().decode("utf-8", "replace")


def check(rc, what):
    """0 -> ok; argument errors surface as AssertionError like the reference's asserts
    (pf:253, :479, :484 ...), CUDA failures as MccnnError."""
    if rc == 0:
        return
    msg = "%s: %s" % (what, last_error())
    if rc == -1:
        raise AssertionError(msg)
    raise MccnnError(msg)


def call(name, *args):
    check(getattr(lib(), name)(*args), name)


def launch_count():
    return int(lib().mccnn_launch_count())


def dpitch(D):
    return (int(D) + 3) & ~3


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Raw device address of a torch CUDA tensor (or None)."""
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())
