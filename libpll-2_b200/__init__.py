"""B200-native phylogenetic likelihood engine: host-side Python view.

The product is ``libpll_b200.so`` (C host layer + sm_100a CUDA kernels,
``csrc/``), a C-ABI drop-in for libpll-2's partition API.  This package only
holds the ctypes binding used by tests and bench.py, the synthetic-input
generator and the multi-GPU site-sharding helper.  Import with
``importlib.import_module("libpll-2_b200")`` (the directory name is not a
Python identifier).
"""
import os

from . import capi

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_DIR = os.path.dirname(PKG_DIR)
LIB_PATH = os.path.join(PKG_DIR, "libpll_b200.so")
ORACLE_PATH = os.path.join(REPO_DIR, "oracle", "libplf_oracle.so")
REF_PATH = os.path.join(REPO_DIR, "oracle", "_ref", "libpll_ref.so")

_lib = None


def load() -> "capi.PllLibrary":
    """Load the CUDA engine.  Raises if the native library is missing: there
    is deliberately no CPU or pure-Python fallback."""
    global _lib
    if _lib is None:
        # $PLL_B200_LIB: another build of the same library (A/B measurements against an earlier round)
        _lib = capi.PllLibrary(os.environ.get("PLL_B200_LIB") or LIB_PATH, cuda=True)
    return _lib
