import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def zvx():
    from zvxload import zvx as pkg
    return pkg


@pytest.fixture(scope="session")
def gguf_path(zvx):
    return zvx.synth.write_model(zvx.synth.default_model_path())


@pytest.fixture(scope="session")
def weights(zvx, gguf_path):
    return zvx.gguf_io.read_gguf(gguf_path)[1]


@pytest.fixture(scope="session")
def ctx(weights):
    """The CUDA context.  No fallback: if libzvx.so is missing or there is no sm_100 device, GPU tests fail loudly."""
    from zerovox_cpp_b200 import capi
    c = capi.Context(weights, device=0)
    yield c
    c.close()


def golden(L):
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", f"ref_L{L}.npz"))
