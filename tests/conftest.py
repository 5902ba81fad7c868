import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


    # RGBL_EMULATE=1 (test infrastructure, never set by the driver): run the -m gpu parity tests WITHOUT a GPU against
    # tests/cuda_emu/build/librgbl_b200_emu.so = the whole library compiled by g++ over the CUDA-on-CPU shim (one OS thread per
    # CUDA thread).  Slow (minutes per test file) but it exercises the real kernels' device code through the real C ABI.
    if os.environ.get("RGBL_EMULATE") == "1":
        import importlib.util
        spec = importlib.util.spec_from_file_location("cuda_emu_build", ROOT / "tests" / "cuda_emu" / "build.py")
        mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
        from orb_slam3_rgbl_b200 import _lib
        _lib.LIB_PATH = mod.build_full()


def _have_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def have_gpu():
    return _have_gpu()


@pytest.fixture(scope="session")
def report_dir():
    d = ROOT / "gpurun_out"
    d.mkdir(exist_ok=True)
    return d
