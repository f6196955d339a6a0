import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# RDN_SIMT_EMU=1: run the `-m gpu` tests on the CPU against tests/simt/_build/librdn_rt_emu.so — the same kernel sources
# compiled by g++ and executed one fiber per CUDA thread (tests/simt/simt_emu.h).  Test infrastructure only: it is switched on
# by tests/test_simt_emulation.py (in a subprocess) and by nobody else; the package itself never looks for that library.
SIMT_EMU = os.environ.get("RDN_SIMT_EMU") == "1"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    if SIMT_EMU:
        sys.path.insert(0, os.path.join(ROOT, "tests", "simt"))
        import build_emu
        import torch_on_host
        from rendiation_b200 import api
        api.LIB_PATH = build_emu.build()  # before anything calls api.lib()
        torch_on_host.install()


def _cuda_available() -> bool:
    if SIMT_EMU:
        return True
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _cuda_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the oracle (gcc) and the product library (nvcc) once per session."""
    import oracle
    oracle.build()
    from rendiation_b200 import build as b
    b.build()
    yield
