import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def sp():
    from __graft_entry__ import load_package, build
    build()
    return load_package()


@pytest.fixture(scope="session")
def orc():
    from oracle import orc as m
    m.build()
    return m


@pytest.fixture(scope="session")
def synth(sp):
    import importlib
    return importlib.import_module("stair_perception_b200.synth")


@pytest.fixture(scope="session")
def samples(sp):
    out = {}
    for name in ("0000", "0001"):
        rec, w, h = sp.read_pcd(os.path.join(GOLDEN, "samples", name + "_cloud.pcd"))
        out[name] = (rec, w, h)
    return out


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
