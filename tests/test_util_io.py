"""CPU: the host-side I/O helpers keep the reference's file formats (util.py:6-75)."""
import importlib
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
util = importlib.import_module("mc-cnn-python_b200.util")


def test_pfm_layout_is_bottom_up_little_endian_scale_minus_one(tmp_path):
    m = np.arange(12, dtype=np.float32).reshape(3, 4) + 0.5
    path = str(tmp_path / "d.pfm")
    util.writePfm(m, path)
    raw = open(path, "rb").read()
    header = b"Pf\n4 3\n-1.0\n"                                   # util.py:61-63
    assert raw.startswith(header) and len(raw) == len(header) + 48
    first = struct.unpack("<4f", raw[len(header):len(header) + 16])
    assert list(first) == list(m[2])                               # bottom row first (util.py:67)
    assert np.array_equal(util.readPfm(path), m)


def test_read_big_endian_pfm(tmp_path):
    m = np.array([[1.0, 2.0], [3.0, 4.0]], dtype=np.float32)
    path = str(tmp_path / "be.pfm")
    with open(path, "wb") as f:
        f.write(b"Pf\n2 2\n1.0\n")
        f.write(np.flipud(m).astype(">f4").tobytes())
    assert np.array_equal(util.readPfm(path), m)


def test_parse_calib_reads_lines_five_to_seven(tmp_path):
    path = str(tmp_path / "calib.txt")
    open(path, "w").write("cam0=[1 0 0; 0 1 0; 0 0 1]\ncam1=[1 0 0; 0 1 0; 0 0 1]\ndoffs=131.111\nbaseline=193.001\n"
                          "width=1500\nheight=1000\nndisp=256\nisint=0\nvmin=31\nvmax=257\n")
    assert util.parseCalib(path) == (1000, 1500, 256)              # (height, width, ndisp), util.py:43


def test_pgm_and_time_file(tmp_path):
    m = np.array([[0.4, 0.5, 1.5, 2.5], [254.6, 300.0, -3.0, np.nan]], dtype=np.float32)
    path = str(tmp_path / "d.pgm")
    util.saveDisparity(m, path)
    raw = open(path, "rb").read()
    assert raw.startswith(b"P5\n4 2\n255\n")
    assert list(raw[-8:]) == [0, 0, 2, 2, 255, 255, 0, 0]          # round half to even, saturate, NaN -> 0
    util.saveTimeFile(1.25, str(tmp_path / "t.txt"))
    assert open(str(tmp_path / "t.txt")).read() == "1.25"
    util.recurMk(str(tmp_path / "a" / "b" / "c"))
    assert os.path.isdir(str(tmp_path / "a" / "b" / "c"))
    g = util.normal(0, 6)
    assert abs(g(0.0) - 1.0 / (np.sqrt(2 * np.pi) * 6)) < 1e-12


def test_pfm_bytes_equal_the_reference_writer(tmp_path):
    """tests/golden/file_formats.npz holds the bytes the reference's own util.writePfm (util.py:54-70, run through
    oracle/gen_golden.py) wrote for a map with inf / NaN / -0.0: the drop-in writer must produce the same file, and
    the drop-in reader the same array the reference's readPfm (util.py:6-25) returned."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "file_formats.npz"))
    path = str(tmp_path / "g.pfm")
    util.writePfm(g["pfm_map"], path)
    assert open(path, "rb").read() == g["pfm_bytes"].tobytes()
    back = util.readPfm(path)
    assert np.array_equal(back, g["pfm_read_back"], equal_nan=True)
    assert np.array_equal(back.view(np.uint32), g["pfm_map"].view(np.uint32))      # bit for bit incl. NaN payload, -0.0
