import os
import sys

# The golden fixtures were written by the reference running on oneDNN's AVX512_CORE bf16 kernels.  Newer hosts (AMX /
# avx512_bf16) pick kernels with another fp32 summation order, which breaks BIT-exactness of the CPU oracle against the
# fixtures (not its correctness).  Pin the ISA before torch / oneDNN initialise so the check is machine-independent.
os.environ.setdefault("ONEDNN_MAX_CPU_ISA", "AVX512_CORE")

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
ORACLE = os.path.join(ROOT, "oracle")
if ORACLE not in sys.path:
    sys.path.insert(0, ORACLE)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
