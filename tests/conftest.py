import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The CUDA library is built in-tree (nvcc cross-compiles without a GPU) before any test touches the package."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "spn_b200_build", os.path.join(ROOT, "conditional-ratspn-for-rl_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if os.path.exists("/usr/local/cuda/bin/nvcc"):
        mod.build()
    yield
