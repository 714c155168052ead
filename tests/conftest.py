import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_package():
    """The product package (its directory name is not an identifier)."""
    import importlib
    return importlib.import_module("mc-cnn-python_b200")


@pytest.fixture(scope="session")
def pkg():
    return load_package()


@pytest.fixture(scope="session")
def pf(pkg):
    return pkg.process_functional


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))


@pytest.fixture(scope="session")
def oracle():
    import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session", params=["pipeline_a", "pipeline_b"])
def pipeline_golden(request):
    return load_golden(request.param)


@pytest.fixture(scope="session")
def integer_golden():
    return load_golden("integer_cases")


@pytest.fixture(scope="session")
def features_golden():
    return load_golden("features_ckpt")


@pytest.fixture(scope="session")
def flat_golden():
    """Two-level image (arms at the 13-pixel limit, regions up to 729) through the reference's own code."""
    return load_golden("flat_regions")


@pytest.fixture(scope="session")
def checkpoint_golden():
    """(weights, biases) of the reference's shipped checkpoint (ten conv tensors, oracle/gen_golden.py)."""
    g = load_golden("checkpoint_tensors")
    return [g["conv%d_weights" % i] for i in range(1, 6)], [g["conv%d_biases" % i] for i in range(1, 6)]
