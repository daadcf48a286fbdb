import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # a gpu-marked test without a device is a skip here; on a box that is SUPPOSED to have a GPU
    # (B200_REQUIRE_GPU=1) it is an error, so a broken driver cannot pass as "all skipped"
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    if os.environ.get("B200_REQUIRE_GPU") and any("gpu" in item.keywords for item in items):
        raise pytest.UsageError("B200_REQUIRE_GPU is set but torch sees no CUDA device")
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def b200env():
    import b200env as mod
    return mod


@pytest.fixture(scope="session")
def sf_graph(b200env):
    return b200env.load_sioux_falls()


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    return load


@pytest.fixture(scope="session")
def sf_oracle(sf_graph):
    from oracle.c_oracle import COracle
    return COracle.from_graph_data(sf_graph)


@pytest.fixture(scope="session")
def native_lib(b200env):
    """Builds (if stale) and loads the CUDA library; nvcc cross-compiles without a GPU."""
    b200env.build_native()
    return b200env._native.lib()
