import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# Oracle (torch fp32 on this host's CPU) vs. fixtures written by the reference on the host that generated them: same ops,
# but the fp32 summation order of torch's CPU convolutions depends on the vector ISA (AVX / AVX2 / AVX-512) and thread
# count, i.e. a few ulp at |logit| ~ 10.  Integer outputs (indices, labels, counts) are still compared exactly.
ORACLE_FIXTURE_TOL = 1e-4


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) GPU; run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def native_lib():
    """libav1p.so, built in-tree.  GPU tests must run on the native path - fail loudly if it is missing."""
    import __graft_entry__ as g
    g.build()
    from cnn_av1_research_b200 import _native
    return _native.lib()


@pytest.fixture(scope="session")
def cuda_device(native_lib):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    return torch.device("cuda:0")
