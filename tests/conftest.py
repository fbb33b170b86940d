import importlib
import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if os.path.join(REPO, "tests") not in sys.path:
    sys.path.insert(0, os.path.join(REPO, "tests"))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

pkg = importlib.import_module("libpll-2_b200")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def b200pkg():
    return pkg


@pytest.fixture(scope="session")
def reflib():
    """The unmodified reference, prebuilt by oracle/Makefile (`make ref`)."""
    if not os.path.exists(pkg.REF_PATH):
        pytest.skip("oracle/_ref/libpll_ref.so not built (needs /root/reference)")
    return pkg.capi.PllLibrary(pkg.REF_PATH, cuda=False)


@pytest.fixture(scope="session")
def oracle():
    import oracle_api

    return oracle_api.load(pkg.ORACLE_PATH)


@pytest.fixture(scope="session")
def cudalib():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return pkg.load()
