import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
if str(ROOT / "tests") not in sys.path:
    sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def excenvs():
    import exciting_environments_b200 as ex

    return ex


@pytest.fixture(scope="session")
def hostcheck_lib():
    """tests/hostcheck: host instantiation of the kernels' __host__ __device__ math (test-only)."""
    import ctypes

    src = ROOT / "tests" / "hostcheck" / "hostcheck.cu"
    out = ROOT / "tests" / "hostcheck" / "_build" / "libhostcheck.so"
    hdrs = list((ROOT / "exciting-environments_b200" / "csrc").glob("*.cuh")) + [src]
    if not out.exists() or any(h.stat().st_mtime > out.stat().st_mtime for h in hdrs):
        out.parent.mkdir(exist_ok=True)
        subprocess.run(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17",
                        "--expt-relaxed-constexpr", "-fmad=false", "-Xcompiler", "-fPIC,-ffp-contract=off", "-w",
                        "-shared", "-o", str(out), str(src)], check=True)
    return ctypes.CDLL(str(out))
