import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__)),
          os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def hexo_lib():
    """The product's C-ABI library, built in-tree if necessary (nvcc, no GPU needed)."""
    from hestonexotics_b200 import _lib, build
    if build.needs_build():
        build.build()
    return _lib.load()


@pytest.fixture(scope="session")
def gpu(hexo_lib):
    """A CUDA device, through the product library only."""
    from hestonexotics_b200 import _lib
    if hexo_lib.hexo_gpu_device_count() < 1:
        pytest.fail("test marked gpu but no CUDA device is visible")
    _lib.check(hexo_lib.hexo_gpu_init(0))
    return hexo_lib
