"""mc-cnn-python_b200: B200-native stereo-matching hot path of Jackie-Chou/MC-CNN-python.

The directory name is not a Python identifier; import it with
``importlib.import_module("mc-cnn-python_b200")`` (see __graft_entry__.py), or put this directory
on ``sys.path`` and ``from process_functional import *`` exactly like the reference's ``src/``.
"""
from . import _ffi                      # noqa: F401
from . import checkpoint                # noqa: F401
from . import process_functional        # noqa: F401
from . import model                     # noqa: F401
from . import pipeline                  # noqa: F401
from .model import NET                  # noqa: F401
from .pipeline import StereoMatcher, match_pair, shard_window, DEFAULTS   # noqa: F401
from . import slab                      # noqa: F401
from .slab import SlabPlan, SlabRank, SlabMatcher, LocalComm, DistComm, run_slabs, run_slabs_p2p   # noqa: F401
