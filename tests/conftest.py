"""pytest configuration: registers the `gpu` marker and puts the repo root on sys.path.

`-m "not gpu"` (run on CPU-only machines): the oracle against the golden fixtures, host
logic, the C-ABI export table, gloo multi-process tests.
`-m gpu` (run on a B200): parity of the CUDA path, called through the C ABI, against the oracle.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); skipped by -m 'not gpu'")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(params=["fp32", "tf32"])
def conv_mode(request):
    """Run a GPU test twice: on the exact fp32 FFMA convolution kernels (strict tolerances: proves the
    algebra, the autograd wiring and the host logic) and on the tcgen05 kind::tf32 kernels (the
    north-star tolerance of 1e-3 per op; see tests/tolerances.py for how gradients through the
    leaky-ReLU mask are compared)."""
    from ideas_b200 import _lib
    from ideas_b200.stylegan2.op import conv as C
    old = C.set_default_impl(_lib.IMPL_SIMT if request.param == "fp32" else _lib.IMPL_AUTO)
    yield request.param
    C.set_default_impl(old)
