"""Host-side I/O helpers with the names and file formats of the reference's ``src/util.py`` (util.py:6-86):
Middlebury-v3 PFM maps (bottom-up rows, little-endian, scale -1.0), ``calib.txt`` parsing, PGM preview images,
the one-number time file, and ``normal`` (the Gaussian used by the bilateral filter).  Written for Python 3
(the reference opens binary files in text mode, which only works under Python 2) and vectorised with NumPy."""
import os
import struct

import numpy as np


def readPfm(filename):
    """util.py:6-25.  Single-channel PFM -> float32 [H, W], top row first."""
    with open(filename, "rb") as f:
        assert f.readline().strip() == b"Pf", "one sample per pixel expected"       # util.py:9
        items = f.readline().strip().split()
        width, height = int(items[0]), int(items[1])
        scale = float(f.readline().strip())
        dtype = "<f4" if scale < 0 else ">f4"                                          # util.py:15-18
        data = np.frombuffer(f.read(4 * width * height), dtype=dtype)
    assert data.size == width * height, "truncated PFM"
    return np.flipud(data.reshape(height, width)).astype(np.float32)                    # rows are stored bottom-up


def writePfm(disparity_map, filename):
    """util.py:54-70.  float32 [H, W] -> PFM: "Pf", "W H", "-1.0", then rows bottom-up, little-endian."""
    assert len(disparity_map.shape) == 2
    height, width = disparity_map.shape
    with open(filename, "wb") as o:
        o.write(b"Pf\n")
        o.write("{} {}\n".format(width, height).encode("ascii"))
        o.write(b"-1.0\n")
        o.write(np.flipud(np.asarray(disparity_map, dtype=np.float32)).astype("<f4").tobytes())


def parseCalib(filename):
    """util.py:27-43.  Lines 5-7 of a Middlebury calib.txt are width=, height=, ndisp=; returns (height, width, ndisp)."""
    with open(filename, "r") as f:
        lines = f.readlines()
    values = []
    for line in lines[4:7]:
        line = line.strip()
        values.append(int(line[line.find("=") + 1:]))
    width, height, ndisp = values
    return height, width, ndisp


def normal(mean, std_dev):
    """util.py:45-48."""
    constant1 = 1. / (np.sqrt(2 * np.pi) * std_dev)
    constant2 = -1. / (2 * std_dev * std_dev)
    return lambda x: constant1 * np.exp(constant2 * ((x - mean) ** 2))


def saveDisparity(disparity_map, filename):
    """util.py:50-52 (cv2.imwrite of a float map to .pgm): 8-bit binary PGM, values rounded to nearest (ties to
    even) and saturated to [0, 255] like OpenCV's saturate_cast; NaN -> 0."""
    assert len(disparity_map.shape) == 2
    height, width = disparity_map.shape
    a = np.nan_to_num(np.asarray(disparity_map, dtype=np.float64), nan=0.0, posinf=255.0, neginf=0.0)
    img = np.clip(np.rint(a), 0, 255).astype(np.uint8)
    with open(filename, "wb") as o:
        o.write("P5\n{} {}\n255\n".format(width, height).encode("ascii"))
        o.write(img.tobytes())


def saveTimeFile(times, path):
    """util.py:72-75."""
    with open(path, "w") as o:
        o.write("{}".format(times))


def testMk(dirName):
    """util.py:77-79."""
    if not os.path.isdir(dirName):
        os.mkdir(dirName)


def recurMk(path):
    """util.py:81-86: create every directory along `path`."""
    os.makedirs(path, exist_ok=True)
