"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Loader that executes the reference's own ``src/process_functional.py`` (Python 2
+ TensorFlow 1.x) under Python 3 *without copying it*: the source is read from
``/root/reference`` at run time, patched in memory and ``exec``-ed.

Patches:
  1. ``print ...`` statements            -> ``pass``
  2. index-valued ``/2``                 -> ``//2``  (pf:22-23, 411-414, 431-432, 442-445)
  3. ``tensorflow`` / ``model`` / ``util`` stubbed in ``sys.modules``
     (``util.normal`` restated from util.py:45-48)

This only works where ``/root/reference`` exists (the authoring container); it is
used by ``oracle/gen_golden.py`` to produce the committed fixtures under
``tests/golden/`` and by ``tests/test_oracle_vs_reference.py`` (skipped when the
reference tree is absent, e.g. on the GPU box).
"""
import os
import re
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("MCCNN_REFERENCE_ROOT", "/root/reference")
_PF_PATH = os.path.join(REFERENCE_ROOT, "src", "process_functional.py")


def reference_available():
    return os.path.isfile(_PF_PATH)


def _patch_source(src):
    out = []
    for line in src.split("\n"):
        m = re.match(r"^(\s*)print\b(?!\s*\()", line)
        if m:
            out.append(m.group(1) + "pass")
            continue
        line = line.replace("(patch_height - 1)/2", "(patch_height - 1)//2")
        line = line.replace("(patch_width - 1)/2", "(patch_width - 1)//2")
        line = line.replace("(filter_height-1)/2", "(filter_height-1)//2")
        line = line.replace("(filter_width-1)/2", "(filter_width-1)//2")
        line = line.replace("(filter_height - 1)/2", "(filter_height - 1)//2")
        line = line.replace("(filter_width - 1)/2", "(filter_width - 1)//2")
        out.append(line)
    return "\n".join(out)


def _util_stub():
    util = types.ModuleType("util")

    def normal(mean, std_dev):  # restated from reference util.py:45-48
        constant1 = 1. / (np.sqrt(2 * np.pi) * std_dev)
        constant2 = -1. / (2 * std_dev * std_dev)
        return lambda x: constant1 * np.exp(constant2 * ((x - mean) ** 2))

    util.normal = normal
    return util


_cached = None


def load_reference_pf():
    """Return a module object holding the reference's process_functional functions."""
    global _cached
    if _cached is not None:
        return _cached
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    with open(_PF_PATH, "r") as f:
        src = _patch_source(f.read())
    saved = {k: sys.modules.get(k) for k in ("tensorflow", "model", "util")}
    tf = types.ModuleType("tensorflow")
    model = types.ModuleType("model")
    model.NET = None
    sys.modules["tensorflow"] = tf
    sys.modules["model"] = model
    sys.modules["util"] = _util_stub()
    try:
        mod = types.ModuleType("reference_process_functional")
        mod.__file__ = _PF_PATH
        exec(compile(src, _PF_PATH, "exec"), mod.__dict__)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _cached = mod
    return mod
