import importlib
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, "tests", "golden")
TINY = dict(image_dim=128, text_dim=48)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def pkg(sub=None):
    name = "recommendar-systems_b200" + ("." + sub if sub else "")
    return importlib.import_module(name)


def golden(tag):
    return np.load(os.path.join(GOLDEN, tag + ".npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def tiny_data():
    return pkg("synth").make_dataset("tiny", **TINY)


@pytest.fixture(scope="session")
def tiny_train(tiny_data):
    return tiny_data.split(0)
