"""CPU: the C-ABI library loads and exports every symbol include/mccnn_b200.h declares, the ctypes
table mirrors the header, and the host-side logic around it (no compute calls without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mccnn_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = re.findall(r"\b(?:int|size_t|const char \*|unsigned long long)\s*\*?\s*(mccnn_\w+)\s*\(([^;{]*)\)\s*;", src)
    return {name: args for name, args in protos}


@pytest.fixture(scope="module")
def built_lib(pkg):
    import importlib
    build = importlib.import_module("mc-cnn-python_b200.build")
    return ctypes.CDLL(build.build())


def test_header_declares_the_hot_path():
    fns = header_functions()
    for name in ["mccnn_features", "mccnn_cost_volume", "mccnn_cross_arms", "mccnn_cbca", "mccnn_sgm_pass",
                 "mccnn_sgm_average", "mccnn_wta", "mccnn_lr_interp", "mccnn_subpixel", "mccnn_median",
                 "mccnn_bilateral", "mccnn_last_error", "mccnn_dhw_to_hwd", "mccnn_hwd_to_dhw"]:
        assert name in fns


def test_library_exports_every_declared_symbol(built_lib):
    for name in header_functions():
        assert hasattr(built_lib, name), "libmccnn_b200.so does not export %s" % name


def test_ctypes_table_mirrors_header(pkg):
    fns = header_functions()
    sig = pkg._ffi.SIGNATURES
    assert set(sig) == set(fns), (set(sig) ^ set(fns))
    for name, args in fns.items():
        args = args.strip()
        n = 0 if args in ("", "void") else len(args.split(","))
        assert len(sig[name][1]) == n, "%s: header has %d parameters, ctypes table %d" % (name, n, len(sig[name][1]))


def test_no_compute_calls_without_gpu_but_pure_queries_work(pkg):
    lib = pkg._ffi.lib()
    assert lib.mccnn_abi_version() == 2
    for D in (1, 2, 3, 4, 11, 192, 400):
        assert lib.mccnn_dpitch(D) == pkg._ffi.dpitch(D) == (D + 3) // 4 * 4
    # two ping-pong activation maps + 32 range flags + per layer 2..5 the TF32 hi/lo split, the FP16 hi/lo split and a flag block
    assert lib.mccnn_features_scratch_bytes(128, 128, 5, 5) == (2 * 136 * 136 * 64 + 32 + 4 * (3 * 9 * 64 * 64 + 16)) * 4
    assert lib.mccnn_features_weights_bytes(5) == 4 * (3 * 9 * 64 * 64 + 16) * 4
    assert lib.mccnn_sgm_scratch_bytes(16, 40, 8) > 0
    # argument validation happens before any CUDA call and reports through the error string
    rc = lib.mccnn_cost_volume(None, None, None, None, 4, 4, 64, 8, None)
    assert rc == -1 and b"null pointer" in lib.mccnn_last_error()
    one = ctypes.c_void_p(16)
    rc = lib.mccnn_cost_volume(one, one, one, one, 4, 9, 64, 8, None)
    assert rc == -1 and b"W >= ndisp + 2" in lib.mccnn_last_error()
    rc = lib.mccnn_sgm_pass(one, one, one, one, 8, 4, 4, 1, 1, 1.0, 2.0, 4.0, 8.0, 0.1, 1, None)
    assert rc == -1 and b"axis-aligned" in lib.mccnn_last_error()
    with pytest.raises(AssertionError):
        pkg._ffi.check(rc, "sgm_pass")


def test_product_fails_loudly_without_cuda(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(pkg._ffi.MccnnError):
        pkg.process_functional.compute_cost_volume(np.zeros((4, 12, 64), np.float32), np.zeros((4, 12, 64), np.float32), 4)


def test_product_never_imports_the_oracle():
    pkgdir = os.path.join(ROOT, "mc-cnn-python_b200")
    for root, _, files in os.walk(pkgdir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f)).read()
                assert "oracle" not in text.replace("the oracle", ""), "%s mentions the oracle" % f


def test_host_side_tables_match_reference_formulas(pkg, oracle):
    pf = pkg.process_functional
    assert np.array_equal(pf.bilateral_table(5, 5, 0, 6), oracle.bilateral_table(5, 5, 0, 6))
    w1, b1 = pf.glorot_uniform_weights(seed=1234)
    w2, b2 = oracle.glorot_uniform_weights(seed=1234)
    assert all(np.array_equal(a, b) for a, b in zip(w1 + b1, w2 + b2))
    assert w1[0].shape == (3, 3, 1, 64) and w1[1].shape == (3, 3, 64, 64)
    assert abs(np.abs(w1[0]).max() - 0.10127) < 2e-4 and abs(np.abs(w1[1]).max() - 0.07217) < 1e-4


def test_checkpoint_reader_on_reference_bundle(pkg, features_golden):
    prefix = "/root/reference/data/tensorboard_log/model_epoch2000.ckpt"
    if not os.path.isfile(prefix + ".index"):
        pytest.skip("reference checkpoint not present on this box")
    ws, bs = pkg.checkpoint.load_mccnn_weights(prefix)
    g = features_golden
    assert np.array_equal(ws[0], g["conv1_weights"]) and np.array_equal(bs[0], g["conv1_biases"])
    assert np.array_equal(bs[4], g["conv5_biases"])
    np.testing.assert_allclose([float(w.astype(np.float64).sum()) for w in ws], g["weight_sums"], rtol=1e-12)


def test_slab_entry_points_validate_before_touching_the_device(pkg):
    """The multi-GPU entry points (one big pair by disparity slab) reject inconsistent partitions on the host."""
    lib = pkg._ffi.lib()
    one = ctypes.c_void_p(16)
    err = lambda: lib.mccnn_last_error()
    # a slab must lie inside [0, ndisp)
    assert lib.mccnn_cost_volume_slab(one, one, one, one, 4, 40, 64, 16, 12, 8, None) == -1 and b"slab" in err()
    # horizontal passes need whole rows; directions come in pairs selected by `which`
    assert lib.mccnn_sgm_passes_slab(one, one, one, one, one, 8, 4, 20, 2, 10, 0, 1.0, 2.0, 4.0, 8.0, 0.1, 1.5, None) == -1
    assert b"whole rows" in err()
    assert lib.mccnn_sgm_passes_slab(one, one, one, one, one, 8, 4, 20, 0, 20, 2, 1.0, 2.0, 4.0, 8.0, 0.1, 1.5, None) == -1
    # scatter tables: bounds must tile the columns (which = 0) or the granules (which = 1)
    bad = (ctypes.c_int * 3)(0, 7, 19)
    dst = (ctypes.c_void_p * 2)(16, 16)
    rc = lib.mccnn_sgm_passes_slab_to(one, one, one, one, one, 8, 4, 20, 0, 20, 0, 1.0, 2.0, 4.0, 8.0, 0.1, 1.5, 2, bad, dst, dst,
                                      0, None)
    assert rc == -1 and b"tile" in err()
    rc = lib.mccnn_sgm_passes_slab_to(one, one, one, one, one, 8, 4, 20, 0, 20, 0, 1.0, 2.0, 4.0, 8.0, 0.1, 1.5, 9, bad, dst, dst,
                                      0, None)
    assert rc == -1 and b"parts" in err()
    rows = (ctypes.c_int * 3)(0, 2, 5)
    rc = lib.mccnn_cbca_to(one, ctypes.c_void_p(32), ctypes.c_void_p(48), one, one, 8, 4, 20, 2, 14, 2, rows, dst, 0, 2, None)
    assert rc == -1 and b"tile" in err()
    rows = (ctypes.c_int * 3)(0, 2, 4)
    rc = lib.mccnn_cbca_to(one, ctypes.c_void_p(32), ctypes.c_void_p(48), one, one, 8, 4, 20, 2, 14, 2, rows, dst, 1, 2, None)
    assert rc == -1 and b"pitch" in err()
    assert lib.mccnn_wta_combine(one, one, one, 2, 3, 2, 2, None) == -1                 # slab stride smaller than a map
    assert lib.mccnn_copy3d(ctypes.c_void_p(8), one, 1, 1, 1, 0, 0, 0, 0, None) == -1 and b"aligned" in err()
