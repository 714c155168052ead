"""GPU: the REFERENCE's own match.py (src/match.py:56-185, py3-patched copy staged in oracle/_ref by
oracle/stage_ref.py) driving the drop-in modules: with mc-cnn-python_b200/ first on sys.path its
`from process_functional import *` (match.py:13) and `import util` (match.py:5) bind the CUDA-backed replacements, and
the ten calls of match.py:132-175 run unchanged.  Its PFM / PGM outputs must equal what the repository's match.py
writes for the same list.  Skipped where oracle/_ref is absent (a checkout that never saw /root/reference)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "mc-cnn-python_b200")
REF_MATCH = os.path.join(ROOT, "oracle", "_ref", "match.py")

RUNNER = r"""
import sys, types, runpy
sys.modules["tensorflow"] = types.ModuleType("tensorflow")      # match.py:9 imports it and never uses it
sys.path.insert(0, %r)                                          # process_functional / util / model = the drop-ins
sys.argv = ["match.py"] + %r
runpy.run_path(%r, run_name="__main__")
"""


def test_reference_match_py_runs_on_the_drop_in_modules(tmp_path):
    cv2 = pytest.importorskip("cv2")
    if not os.path.isfile(REF_MATCH):
        pytest.skip("oracle/_ref/match.py not staged (python oracle/stage_ref.py needs /root/reference)")
    data = tmp_path / "data"
    H, W, D = 40, 72, 16
    rng = np.random.default_rng(5)
    paths = []
    for name in ("A", "B"):
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
    outs = {}
    for who, script in (("reference", REF_MATCH), ("repo", os.path.join(PKG, "match.py"))):
        save = tmp_path / ("out_" + who)
        save.mkdir()
        # integer-valued hyper-parameters are given explicitly: the reference declares them type=float
        # (match.py:34-36) and then uses them in range() / as array sizes, which only int-valued ints survive
        argv = ["--list_file", str(lst), "--data_dir", str(data), "--save_dir", str(save), "-t", "x", "-s", "0", "-e", "1"]
        env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "0"))
        r = subprocess.run([sys.executable, "-c", RUNNER % (PKG, argv, script)], stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, text=True, timeout=600, env=env, cwd=str(tmp_path))
        assert r.returncode == 0, r.stdout[-3000:]
        outs[who] = save
    util = __import__("importlib").import_module("mc-cnn-python_b200.util")
    for name in ("A", "B"):
        a = outs["reference"] / "submit_x" / name
        b = outs["repo"] / "submit_x" / name
        pa, pb = (a / "disp0MCCNN.pfm").read_bytes(), (b / "disp0MCCNN.pfm").read_bytes()
        assert pa == pb, "PFM of %s differs between the reference's match.py and the repository's" % name
        d = util.readPfm(str(a / "disp0MCCNN.pfm"))
        assert d.shape == (H, W) and np.isfinite(d).mean() > 0.9
        ga = (outs["reference"] / "submit_x_imgs" / name / "disp0MCCNN.pgm").read_bytes()
        gb = (outs["repo"] / "submit_x_imgs" / name / "disp0MCCNN.pgm").read_bytes()
        assert ga == gb
        assert float((a / "timeMCCNN.txt").read_text()) > 0
