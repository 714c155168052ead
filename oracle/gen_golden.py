"""TEST INFRASTRUCTURE: generates tests/golden/*.npz by running the REFERENCE ITSELF.

Run in the authoring container (needs /root/reference):

    python oracle/gen_golden.py

Post-CNN stages: /root/reference/src/process_functional.py executed through
oracle/ref_loader.py (py2->py3 shim, NumPy 2.3.5) on seeded synthetic inputs.
CNN: TensorFlow is not installable here, so the feature golden is produced by a torch-CPU
float64 conv2d restatement of model.py:40-64 on the reference's SHIPPED checkpoint (read with the
product's TF-bundle reader, CRC32C-verified); it pins the C oracle and the checkpoint reader, not
TensorFlow's own arithmetic ("parity unpinned" for TF itself).

The fixtures are small (a few hundred KB in total) and committed; tests never read
/root/reference.
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
from ref_loader import load_reference_pf, REFERENCE_ROOT  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def synth_inputs(seed, H, W, D, levels, shift):
    """Seeded inputs: unit-norm N(0,1) features; blurred-noise image quantised to `levels` grey levels
    then (x-mean)/std as match.py:118-123; right image = left shifted by `shift` px."""
    import cv2
    rng = np.random.default_rng(seed)
    fl = rng.standard_normal((H, W, 64)).astype(np.float32)
    fl /= np.linalg.norm(fl, axis=-1, keepdims=True)
    fr = rng.standard_normal((H, W, 64)).astype(np.float32)
    fr /= np.linalg.norm(fr, axis=-1, keepdims=True)
    base = cv2.GaussianBlur(rng.random((H, W)).astype(np.float32), (0, 0), 2.0)
    q = np.floor((base - base.min()) / (np.ptp(base) + 1e-9) * levels).astype(np.float32)
    qr = np.roll(q, -shift, axis=1)
    li = ((q - np.mean(q, axis=(0, 1))) / np.std(q, axis=(0, 1)))[..., None].astype(np.float32)
    ri = ((qr - np.mean(qr, axis=(0, 1))) / np.std(qr, axis=(0, 1)))[..., None].astype(np.float32)
    return fl, fr, li, ri


def gen_pipeline(pf, name, seed, H, W, D, levels, shift):
    fl, fr, li, ri = synth_inputs(seed, H, W, D, levels, shift)
    g = dict(fl=fl, fr=fr, left_image=li, right_image=ri, ndisp=np.int32(D))
    L, R = pf.compute_cost_volume(fl, fr, D)
    g["cv_L"], g["cv_R"] = L, R
    reg, num = pf.compute_cross_region(li, 0.02, 14)
    g["region_num_left"] = num
    # the explicit list is big ([H,W,784,2]); keep a checksum-friendly compressed copy of one row band
    g["region_left_rows0_4"] = reg[:4].astype(np.int16)
    L1, R1 = pf.cost_volume_aggregation(li, ri, L, R, 0.02, 14, 2)
    g["cbca1_L"], g["cbca1_R"] = L1, R1
    # single in-place passes, each from the CBCA output (not chained), for per-direction gating
    for r in [(0, 1), (0, -1), (-1, 0), (1, 0)]:
        p1 = 2.3 if r[0] == 0 else 2.3 / 1.5
        for ch, src in (("L", L1), ("R", R1)):
            x = src.copy()
            y = pf.semi_global_matching(li, ri, x, r, p1, 55.9, 4, 8, 0.08, ch)
            assert y is x
            g["sgm_%s_%d_%d" % (ch, r[0], r[1])] = y
    Ls, Rs = pf.SGM_average(L1.copy(), R1.copy(), li, ri, 2.3, 55.9, 4, 8, 0.08, 1.5)
    g["sgm_L"], g["sgm_R"] = Ls, Rs
    L2, R2 = pf.cost_volume_aggregation(li, ri, Ls, Rs, 0.02, 14, 16)
    g["cbca2_L"], g["cbca2_R"] = L2, R2
    dl, dr = pf.disparity_prediction(L2, R2)
    g["wta_L"], g["wta_R"] = dl, dr
    d = pf.interpolation(dl, dr, D)
    g["interp"] = d
    d = pf.subpixel_enhance(d, L2)
    g["subpixel"] = d
    d = pf.median_filter(d, 5, 5)
    g["median"] = d
    d = pf.bilateral_filter(li, d, 5, 5, 0, 6, 2)
    g["bilateral"] = d
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **g)
    print(name, "max region", int(num.max()), "labels via interp ok; bytes",
          os.path.getsize(os.path.join(GOLDEN, name + ".npz")))


def gen_integer_cases(pf):
    """Integer-valued float32 costs: WTA with ties (bit-exact gate of north_star) and an SGM case
    that is exact in float32 (integer costs, dyadic penalties P1=2, P2=56, Q=4,8)."""
    rng = np.random.default_rng(7)
    D, H, W = 12, 14, 30
    L = rng.integers(0, 8, (D, H, W)).astype(np.float32)       # many ties
    R = rng.integers(0, 256, (D, H, W)).astype(np.float32)
    dl, dr = pf.disparity_prediction(L, R)
    g = dict(L=L, R=R, wta_L=dl, wta_R=dr)
    _, _, li, ri = synth_inputs(11, H, W, D, 5, 3)
    g["left_image"], g["right_image"] = li, ri
    for r in [(0, 1), (0, -1), (-1, 0), (1, 0)]:
        for ch, src in (("L", R), ("R", R)):
            x = src.copy()
            pf.semi_global_matching(li, ri, x, r, 2.0, 56.0, 4, 8, 0.08, ch)
            g["sgm_int_%s_%d_%d" % (ch, r[0], r[1])] = x
    # interpolation on random integer disparity maps (all three labels, border cases)
    rl = rng.integers(0, D, (H, W)).astype(np.float32)
    rr = rng.integers(0, D, (H, W)).astype(np.float32)
    g["rand_dl"], g["rand_dr"] = rl, rr
    g["rand_interp"] = pf.interpolation(rl, rr, D)
    # sub-pixel on half-integer disparities incl. the int() truncation corner (d = 0.5, d = D-1.5)
    hd = (rng.integers(0, 2 * D - 1, (H, W)).astype(np.float32)) / 2.0
    vol = rng.standard_normal((D, H, W)).astype(np.float32)
    g["half_disp"], g["sub_vol"] = hd, vol
    with np.errstate(all="ignore"):
        g["half_subpixel"] = pf.subpixel_enhance(hd, vol)
        g["int_subpixel"] = pf.subpixel_enhance(rl, L)       # integer costs: 0/0 and x/0 appear
    g["half_median"] = pf.median_filter(g["half_subpixel"], 5, 5)
    g["half_bilateral"] = pf.bilateral_filter(li, g["half_median"], 5, 5, 0, 6, 2)
    np.savez_compressed(os.path.join(GOLDEN, "integer_cases.npz"), **g)
    print("integer_cases bytes", os.path.getsize(os.path.join(GOLDEN, "integer_cases.npz")))


def gen_features():
    import torch
    import torch.nn.functional as F
    spec = importlib.util.spec_from_file_location(
        "mccnn_ckpt", os.path.join(ROOT, "mc-cnn-python_b200", "checkpoint.py"))
    ck = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ck)
    prefix = os.path.join(REFERENCE_ROOT, "data", "tensorboard_log", "model_epoch2000.ckpt")
    ws, bs = ck.load_mccnn_weights(prefix)
    rng = np.random.default_rng(3)
    H, W = 20, 27
    img = rng.standard_normal((H, W)).astype(np.float32)

    def torch_net(img, ws, bs):
        H, W = img.shape
        x = torch.zeros(1, 1, H + 10, W + 10, dtype=torch.float64)
        x[0, 0, 5:5 + H, 5:5 + W] = torch.from_numpy(img).double()
        for i, (w, b) in enumerate(zip(ws, bs)):
            x = F.conv2d(x, torch.from_numpy(w).permute(3, 2, 0, 1).double(), torch.from_numpy(b).double())
            if i < len(ws) - 1:
                x = torch.relu(x)
            x = x.float().double()          # layer outputs are stored in float32, as TF does
        x = x[0].permute(1, 2, 0)
        x = x * torch.rsqrt(torch.clamp((x * x).sum(-1, keepdim=True), min=1e-12))
        return x.float().numpy()

    g = dict(image=img, features=torch_net(img, ws, bs))
    # same image through seeded glorot-uniform weights (what BASELINE config 1 "random-init" uses);
    # this one is checkable on the GPU box, where the reference checkpoint does not exist
    import oracle as O
    gw, gb = O.glorot_uniform_weights(seed=1234)
    g["glorot_seed"] = np.int32(1234)
    g["features_glorot"] = torch_net(img, gw, gb)
    # a few weight known-answers so the checkpoint reader is pinned without shipping the checkpoint
    g["conv1_weights"] = ws[0]
    g["conv1_biases"] = bs[0]
    g["conv5_biases"] = bs[4]
    g["weight_sums"] = np.array([float(w.astype(np.float64).sum()) for w in ws])
    # (the full checkpoint is NOT copied into the repo; tests that need it read it from
    #  /root/reference when present and otherwise use seeded glorot weights)
    np.savez_compressed(os.path.join(GOLDEN, "features_ckpt.npz"), **g)
    print("features_ckpt bytes", os.path.getsize(os.path.join(GOLDEN, "features_ckpt.npz")))


def two_level_images(H, W, shift):
    """Piece-wise constant pair: two grey levels split by a wavy vertical boundary, so that arms run to the
    distance limit (13 pixels) almost everywhere and regions reach 27 x 27 = 729 (pf:585-599, :612-626)."""
    hh, ww = np.mgrid[0:H, 0:W + shift]
    q = (ww > (W + shift) / 2 + 4 * np.sin(hh / 3.0)).astype(np.float32) * 200.0 + 20.0
    left, right = q[:, :W], q[:, shift:shift + W]
    li = ((left - np.mean(left, axis=(0, 1))) / np.std(left, axis=(0, 1)))[..., None].astype(np.float32)
    ri = ((right - np.mean(right, axis=(0, 1))) / np.std(right, axis=(0, 1)))[..., None].astype(np.float32)
    return li, ri


def gen_flat(pf):
    """Worst-case cross regions (the other fixtures top out at 76 pixels): counts, the explicit list of one row band,
    and the aggregation after 1, 2 and 5 rounds of a random volume, all from the reference's own code."""
    H, W, D = 30, 64, 6
    li, ri = two_level_images(H, W, 3)
    rng = np.random.default_rng(404)
    L = rng.standard_normal((D, H, W)).astype(np.float32)
    R = rng.standard_normal((D, H, W)).astype(np.float32)
    g = dict(left_image=li, right_image=ri, cv_L=L, cv_R=R)
    reg, num = pf.compute_cross_region(li, 0.02, 14)
    g["region_num_left"] = num
    g["region_left_rows13_15"] = reg[13:15].astype(np.int16)
    _, g["region_num_right"] = pf.compute_cross_region(ri, 0.02, 14)
    for iters in (1, 2, 5):
        a, b = pf.cost_volume_aggregation(li, ri, L, R, 0.02, 14, iters)
        g["cbca%d_L" % iters], g["cbca%d_R" % iters] = a, b
    np.savez_compressed(os.path.join(GOLDEN, "flat_regions.npz"), **g)
    print("flat_regions: max region", int(num.max()), "mean", float(num.mean()), "bytes",
          os.path.getsize(os.path.join(GOLDEN, "flat_regions.npz")))


def gen_checkpoint_tensors():
    """The ten conv tensors of the reference's shipped checkpoint (data/tensorboard_log/model_epoch2000.ckpt, read with
    the product's CRC32C-verified bundle reader), so that real-weight features can be run on the GPU box, where
    /root/reference does not exist (pf:32, :43), plus the float64 torch restatement's features of one image."""
    spec = importlib.util.spec_from_file_location(
        "mccnn_ckpt", os.path.join(ROOT, "mc-cnn-python_b200", "checkpoint.py"))
    ck = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ck)
    prefix = os.path.join(REFERENCE_ROOT, "data", "tensorboard_log", "model_epoch2000.ckpt")
    ws, bs = ck.load_mccnn_weights(prefix)
    g = {}
    for i, (w, b) in enumerate(zip(ws, bs)):
        g["conv%d_weights" % (i + 1)] = w
        g["conv%d_biases" % (i + 1)] = b
    np.savez_compressed(os.path.join(GOLDEN, "checkpoint_tensors.npz"), **g)
    print("checkpoint_tensors bytes", os.path.getsize(os.path.join(GOLDEN, "checkpoint_tensors.npz")))


def gen_file_formats():
    """Bytes written by the reference's own util.py (py3-patched as oracle/stage_ref.py does: binary mode, header as
    bytes): writePfm (util.py:54-70) of a small map incl. inf / NaN / negative values."""
    import tempfile
    import types
    sys.path.insert(0, HERE)
    import stage_ref
    src = stage_ref._patch_util(open(os.path.join(REFERENCE_ROOT, "src", "util.py")).read())
    util = types.ModuleType("reference_util")
    exec(compile(src, "reference util.py", "exec"), util.__dict__)
    rng = np.random.default_rng(9)
    d = (rng.standard_normal((5, 7)) * 40).astype(np.float32)
    d[0, 0], d[1, 2], d[4, 6] = np.inf, np.nan, -0.0
    path = tempfile.mktemp(suffix=".pfm")
    util.writePfm(d, path)
    raw = np.frombuffer(open(path, "rb").read(), dtype=np.uint8)
    back = util.readPfm(path)
    os.remove(path)
    np.savez_compressed(os.path.join(GOLDEN, "file_formats.npz"), pfm_map=d, pfm_bytes=raw, pfm_read_back=back)
    print("file_formats bytes", os.path.getsize(os.path.join(GOLDEN, "file_formats.npz")))


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    pf = load_reference_pf()
    if "--new" in sys.argv:          # only the fixtures added in round 2 (the others are unchanged)
        gen_flat(pf)
        gen_checkpoint_tensors()
        gen_file_formats()
        return
    # few grey levels -> long arms (large cross regions); many levels -> short arms
    gen_pipeline(pf, "pipeline_a", seed=101, H=20, W=40, D=8, levels=5, shift=2)
    gen_pipeline(pf, "pipeline_b", seed=202, H=33, W=29, D=11, levels=40, shift=3)
    gen_integer_cases(pf)
    gen_features()
    gen_flat(pf)
    gen_checkpoint_tensors()
    gen_file_formats()


if __name__ == "__main__":
    main()
