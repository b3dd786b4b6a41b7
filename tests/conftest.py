import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    """The product package (directory name starts with a digit, hence importlib)."""
    return importlib.import_module("3dssd_b200")


@pytest.fixture(scope="session")
def oracle_ops():
    from oracle import ops
    ops.build()
    return ops


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(scope="session")
def ref_ops(cuda):
    """The reference's own kernels (oracle/_ref/libref_ops.so, built from /root/reference by `make -C oracle ref`)."""
    from oracle import ref_ops
    if not ref_ops.available():
        pytest.skip("oracle/_ref/libref_ops.so not built")
    return ref_ops
