import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
# The host framework (Gymnasium) the engine plugs into: the offline install of the reference under baseline/_ref travels
# with the repo snapshot.  It must be importable BEFORE gymnasium_b200 is first imported so the engine subclasses the
# real gymnasium.vector.VectorEnv; B200ENV_FORCE_COMPAT=1 exercises the stand-in types instead.
_REF = os.path.join(ROOT, "baseline", "_ref")
if os.path.isdir(os.path.join(_REF, "gymnasium")) and _REF not in sys.path and not os.environ.get("B200ENV_FORCE_COMPAT"):
    sys.path.insert(0, _REF)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def golden_files(prefix):
    return sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def have_cuda():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def fixture_kwargs(name):
    """Constructor kwargs encoded in a golden fixture's file name (see tests/golden/make_golden.py)."""
    kw = {}
    if "samestep" in name:
        kw["autoreset_mode"] = "SameStep"
    if "sutton" in name:
        kw["sutton_barto_reward"] = True
    if name.startswith("frozenlake"):
        kw["map_name"] = "4x4" if "4x4" in name else "8x8"
        if "noslip" in name:
            kw["is_slippery"] = False
    return kw


def fixture_options(name):
    return {"low": -0.1, "high": 0.1} if "bounds" in name else None
