#!/usr/bin/env python
"""Drop-in for the reference's ``src/match.py`` (inference driver): same flags (match.py:15-43), same file naming
(match.py:46-54) and per-pair stage order (match.py:131-175), Python 3, with the whole pair processed on one B200
through ``pipeline.StereoMatcher`` (device resident, one H2D of the two images and one D2H of the final map).

Parallel runs: the reference shards a list by hand with ``-s/-e`` windows per process (match.py:26-28, :85-90).
Under ``torchrun`` (RANK / WORLD_SIZE / LOCAL_RANK set) this script does the same automatically: the window
[-s, -e] is split into contiguous per-rank windows and each rank uses GPU LOCAL_RANK; no collective is needed.

    python match.py --list_file L --resume CKPT --data_dir D --save_dir S -t tag -s 0 -e 14
"""
import argparse
import os
import sys
import time
from datetime import datetime

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)

import util                                            # noqa: E402  (this directory's util.py)

left_image_suffix = "im0.png"
right_image_suffix = "im1.png"
calib_suffix = "calib.txt"
out_file = "disp0MCCNN.pfm"
out_img_file = "disp0MCCNN.pgm"
out_time_file = "timeMCCNN.txt"


def build_parser():
    p = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter,
                                description="stereo matching based on trained model and post-processing (B200)")
    p.add_argument("-g", "--gpu", type=str, default="0", help="gpu id to use (ignored under torchrun: LOCAL_RANK is used)")
    p.add_argument("-ps", "--patch_size", type=int, default=11, help="length for height/width of square patch")
    p.add_argument("--list_file", type=str, required=True, help="path to file containing left image list")
    p.add_argument("--resume", type=str, default=None, help="TF checkpoint prefix; None = seeded glorot-uniform init")
    p.add_argument("--data_dir", type=str, required=True, help="path to root dir to data.")
    p.add_argument("--save_dir", type=str, required=True, help="path to root dir to save results")
    p.add_argument("-t", "--tag", type=str, required=True, help="tag used to indicate one run")
    p.add_argument("-s", "--start", type=int, required=True, help="index of first image to do matching")
    p.add_argument("-e", "--end", type=int, required=True, help="index of last image to do matching (inclusive)")
    # hyper-parameters (match.py:32-43).  The reference declares the integer ones as float, which breaks range() when
    # they are given on the command line; integer-valued floats are accepted and cast here.
    p.add_argument("--cbca_intensity", type=float, default=0.02)
    p.add_argument("--cbca_distance", type=float, default=14)
    p.add_argument("--cbca_num_iterations1", type=float, default=2)
    p.add_argument("--cbca_num_iterations2", type=float, default=16)
    p.add_argument("--sgm_P1", type=float, default=2.3)
    p.add_argument("--sgm_P2", type=float, default=55.9)
    p.add_argument("--sgm_Q1", type=float, default=4)
    p.add_argument("--sgm_Q2", type=float, default=8)
    p.add_argument("--sgm_D", type=float, default=0.08)
    p.add_argument("--sgm_V", type=float, default=1.5)
    p.add_argument("--blur_sigma", type=float, default=6)
    p.add_argument("--blur_threshold", type=float, default=2)
    # not in the reference: under torchrun, share EVERY pair among all ranks by disparity slab (one big pair over
    # several GPUs, slab.py) instead of giving each rank its own window of pairs
    p.add_argument("--slab", action="store_true")
    return p


def _as_int(name, v):
    assert float(v) == int(v), "--%s must be integer valued" % name
    return int(v)


def read_normalised(path):
    """match.py:118-123: 8-bit grey image -> float32, zero mean / unit (population) std, [H, W, 1]."""
    import cv2
    img = cv2.imread(path, cv2.IMREAD_GRAYSCALE)
    assert img is not None, "cannot read %s" % path
    img = img.astype(np.float32)
    img = (img - np.mean(img, axis=(0, 1))) / np.std(img, axis=(0, 1))
    return np.expand_dims(img, axis=2)


def main(argv=None):
    args = build_parser().parse_args(argv)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if "LOCAL_RANK" not in os.environ:
        os.environ.setdefault("CUDA_VISIBLE_DEVICES", args.gpu)                    # match.py:59
    import torch
    import pipeline                                                                # this directory's pipeline.py
    if "LOCAL_RANK" in os.environ:
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))

    save_res_dir = os.path.join(args.save_dir, "submit_{}".format(args.tag))       # match.py:67-70
    save_img_dir = os.path.join(args.save_dir, "submit_{}_imgs".format(args.tag))
    util.recurMk(save_res_dir)
    util.recurMk(save_img_dir)
    with open(args.list_file, "r") as f:
        img_paths = [line.strip() for line in f.readlines() if line.strip()]

    # the reference's window is [start, end] inclusive (match.py:85-90); split it across ranks
    first, last = max(args.start, 0), min(args.end, len(img_paths) - 1)
    slab_mode = bool(args.slab) and world > 1
    if slab_mode:
        import torch.distributed as dist
        import slab                                                                # this directory's slab.py
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        own_group = not dist.is_initialized()
        if own_group:
            dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
        lo, hi = 0, max(last - first + 1, 0)                                       # every rank works on every pair
    else:
        lo, hi = pipeline.shard_window(max(last - first + 1, 0), rank, world)
    hp = dict(patch_size=args.patch_size, cbca_intensity=args.cbca_intensity,
              cbca_distance=_as_int("cbca_distance", args.cbca_distance),
              cbca_num_iterations1=_as_int("cbca_num_iterations1", args.cbca_num_iterations1),
              cbca_num_iterations2=_as_int("cbca_num_iterations2", args.cbca_num_iterations2),
              sgm_P1=args.sgm_P1, sgm_P2=args.sgm_P2, sgm_Q1=args.sgm_Q1, sgm_Q2=args.sgm_Q2, sgm_D=args.sgm_D,
              sgm_V=args.sgm_V, blur_sigma=args.blur_sigma, blur_threshold=args.blur_threshold)
    matchers = {}
    done = []
    for index in range(first + lo, first + hi):
        left_path = img_paths[index]
        right_path = left_path.replace(left_image_suffix, right_image_suffix)      # match.py:95-97
        calib_path = left_path.replace(left_image_suffix, calib_suffix)
        res_dir = left_path.replace(args.data_dir, save_res_dir)                   # match.py:100-104
        img_dir = left_path.replace(args.data_dir, save_img_dir)
        res_dir = res_dir[:res_dir.rfind(left_image_suffix) - 1]
        img_dir = img_dir[:img_dir.rfind(left_image_suffix) - 1]
        util.recurMk(res_dir)
        util.recurMk(img_dir)
        height, width, ndisp = util.parseCalib(calib_path)
        left_image, right_image = read_normalised(left_path), read_normalised(right_path)
        assert left_image.shape == (height, width, 1)                              # match.py:124-125
        assert right_image.shape == (height, width, 1)
        key = (height, width, ndisp)
        if key not in matchers:
            if slab_mode:
                matchers[key] = slab.SlabMatcher(height, width, ndisp, checkpoint=args.resume, **hp)
            else:
                matchers[key] = pipeline.StereoMatcher(height, width, ndisp, checkpoint=args.resume, **hp)
        torch.cuda.synchronize()
        st = time.time()                                                           # match.py:129
        disparity = matchers[key].run_host(left_image, right_image)                # match.py:131-175
        elapsed = time.time() - st                                                 # match.py:179
        if slab_mode and rank != 0:
            continue                                                               # rank 0 writes the shared pair
        util.saveDisparity(disparity, os.path.join(img_dir, out_img_file))         # match.py:182-184
        util.writePfm(disparity, os.path.join(res_dir, out_file))
        util.saveTimeFile(elapsed, os.path.join(res_dir, out_time_file))
        print("{}: rank {} pair {} ({}x{}x{}) matched in {:.4f} s -> {}".format(datetime.now(), rank, index, height, width,
                                                                                ndisp, elapsed, res_dir))
        done.append(index)
    if slab_mode and own_group:
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
    return done


if __name__ == "__main__":
    main()
