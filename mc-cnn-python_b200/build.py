"""Builds libmccnn_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python mc-cnn-python_b200/build.py [--force] [--verbose]

nvcc cross-compiles without a GPU.  The shared object is git-ignored but travels to the GPU box
with the repo snapshot.  cudart is linked statically so the library only depends on libcuda.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libmccnn_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
          "--expt-relaxed-constexpr", "-DMCCNN_BUILDING"]
# sources whose float32 operations must stay separately rounded (bit-exact stages): no FMA contraction
EXACT = {"refine.cu", "cbca.cu", "sgm.cu"}
SOURCES = ["common.cu", "features.cu", "cost_volume.cu", "cbca.cu", "sgm.cu", "refine.cu"]


def nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.isfile(exe):
        raise RuntimeError("nvcc not found")
    return exe


def _deps_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in os.listdir(root):
            m = max(m, os.path.getmtime(os.path.join(root, f)))
    return max(m, os.path.getmtime(__file__))


def build(force=False, verbose=False):
    if not force and os.path.isfile(LIB) and os.path.getmtime(LIB) >= _deps_mtime():
        return LIB
    os.makedirs(BUILD, exist_ok=True)
    cc = nvcc()

    def compile_one(src):
        obj = os.path.join(BUILD, src.replace(".cu", ".o"))
        cmd = [cc] + ARCH + COMMON + (["-fmad=false"] if src in EXACT else []) + \
              (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stdout))
        return obj, r.stdout

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        results = list(ex.map(compile_one, SOURCES))
    if verbose:
        for _, out in results:
            sys.stdout.write(out)
    objs = [o for o, _ in results]
    cmd = [cc] + ARCH + ["-shared", "-cudart", "static", "-o", LIB] + objs
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s" % r.stdout)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
