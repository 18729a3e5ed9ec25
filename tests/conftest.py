import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _build_native():
    """The CPU oracle (always) and, if a fresh checkout has not been built yet, the sm_100a
    library (nvcc cross-compiles without a GPU; built artefacts are git-ignored)."""
    import shutil
    import subprocess
    from oracle import gsm_oracle
    gsm_oracle.build()
    from gs_marl_b200 import abi
    if not os.path.exists(abi.LIB_PATH) and (shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc")):
        subprocess.run(["make", "-C", os.path.dirname(abi.LIB_PATH), "-j4"], check=True, capture_output=True)
    yield
