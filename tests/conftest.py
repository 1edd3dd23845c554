import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # the schedule code applies np.sqrt to torch tensors exactly like the reference (sampler.py:215-229) so that the
    # float32/float64 promotion is bit-identical; numpy 2 warns about torch's __array_wrap__ signature on every call
    config.addinivalue_line("filterwarnings", "ignore:__array_wrap__ must accept context:DeprecationWarning")


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The CUDA extension is built in-tree (git-ignored); build it once if this checkout has none.
    nvcc cross-compiles for sm_100a without a GPU."""
    from cap4d_b200 import _lib, build

    if not os.path.exists(_lib.LIB_PATH):
        build.build(verbose=False)
