"""pytest configuration: markers + shared fixtures (oracle handle, product library handle)."""
import ctypes
import glob
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def find_openblas():
    """scipy's bundled OpenBLAS: the real LAPACK ?posvx / CBLAS behind the oracle."""
    import scipy
    libs = os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs")
    hits = sorted(glob.glob(os.path.join(libs, "libscipy_openblas*.so")))
    if not hits:
        raise RuntimeError("scipy's libscipy_openblas*.so not found")
    return hits[0]


def load_oracle():
    """Build (if needed) and dlopen oracle/liblm_oracle.so.  TEST INFRASTRUCTURE ONLY."""
    so = os.path.join(ROOT, "oracle", "liblm_oracle.so")
    srcs = [os.path.join(ROOT, "oracle", f) for f in ("lm_oracle.cpp", "models_oracle.cpp")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    lib = ctypes.CDLL(so, mode=ctypes.RTLD_LOCAL)
    lib.oracle_init.argtypes = [ctypes.c_char_p]
    lib.oracle_init.restype = ctypes.c_int
    rc = lib.oracle_init(find_openblas().encode())
    if rc != 0:
        raise RuntimeError(f"oracle_init failed ({rc})")
    return lib


@pytest.fixture(scope="session")
def oracle_lib():
    return load_oracle()


@pytest.fixture(scope="session")
def oracle(oracle_lib):
    from mir_optim_b200.api import ReferenceAPI
    return ReferenceAPI(oracle_lib)
