#!/bin/bash
# On a box with >= 2 GPUs: match.py on a synthetic Middlebury-style list, once as a single process and once under
# torchrun with --slab (every pair shared by both GPUs by disparity slab); the PFM outputs must be identical.
set -eu
T=/tmp/match_slab_check
rm -rf $T; mkdir -p $T
python - <<PY
import numpy as np, cv2, os
rng = np.random.default_rng(5)
H, W, D = 96, 200, 48
paths = []
for name in ("A", "B"):
    d = "$T/data/" + name
    os.makedirs(d)
    base = cv2.GaussianBlur(rng.integers(0, 256, (H, W + 6)).astype(np.uint8), (5, 5), 1.0)
    cv2.imwrite(d + "/im0.png", base[:, :W]); cv2.imwrite(d + "/im1.png", base[:, 6:])
    open(d + "/calib.txt", "w").write("cam0=[]\ncam1=[]\ndoffs=0\nbaseline=1\nwidth=%d\nheight=%d\nndisp=%d\n" % (W, H, D))
    paths.append(d + "/im0.png")
open("$T/list.txt", "w").write("\n".join(paths) + "\n")
PY
mkdir -p $T/one $T/slab
python mc-cnn-python_b200/match.py --list_file $T/list.txt --data_dir $T/data --save_dir $T/one -t x -s 0 -e 1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    mc-cnn-python_b200/match.py --list_file $T/list.txt --data_dir $T/data --save_dir $T/slab -t x -s 0 -e 1 --slab 2>&1 | grep -v -i "warning\|OMP_NUM\|\*\*\*" | tail -3
python - <<PY
import sys, numpy as np
sys.path.insert(0, "mc-cnn-python_b200")
import util
ok = True
for name in ("A", "B"):
    a = util.readPfm("$T/one/submit_x/%s/disp0MCCNN.pfm" % name); b = util.readPfm("$T/slab/submit_x/%s/disp0MCCNN.pfm" % name)
    same = bool(np.array_equal(a, b, equal_nan=True)); ok = ok and same
    print("pair %s: slab PFM == single-process PFM: %s" % (name, same))
sys.exit(0 if ok else 1)
PY
