import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = os.path.join(ROOT, "scouting-ipp_b200")
if PKG not in sys.path:
    sys.path.insert(0, PKG)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    """The reference's own fixtures: 10 label maps + cfg incl. golden PRM* path costs (tests/golden/make_golden.py)."""
    import json

    import numpy as np

    maps = np.load(os.path.join(ROOT, "tests", "golden", "ref_maps.npz"))
    with open(os.path.join(ROOT, "tests", "golden", "ref_maps.json")) as fh:
        cfg = json.load(fh)
    return {name: (maps[name], cfg[name]) for name in sorted(cfg)}
