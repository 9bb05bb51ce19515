import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import __graft_entry__ as entry  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    return entry.load_package()


@pytest.fixture(scope="session")
def oracle(pkg):
    """CPU oracle driven through the same harness (tests only)."""
    return pkg.SpaSM(entry.build_oracle())


@pytest.fixture(scope="session")
def product_lib():
    return entry.build_product()


@pytest.fixture(scope="session")
def gpu(pkg, product_lib):
    """The product: CUDA library through the C ABI.  No fallback."""
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    api = pkg.SpaSM(product_lib)
    assert api.backend.startswith("cuda")
    return api
