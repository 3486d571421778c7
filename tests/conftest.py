"""pytest config: registers the ``gpu`` marker and puts the product package
(``compatibility-family-learning_b200/`` -- not an importable name, so it is a path entry
holding the ``cfl`` package) and the repo root (for ``oracle``) on sys.path."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "compatibility-family-learning_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def pytest_collection_modifyitems(config, items):
    import pytest
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


import pytest  # noqa: E402


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """Builds libcfl_b200.so in-tree when it is missing or stale (nvcc cross-compiles without
    a GPU; on the GPU box the prebuilt .so travels with the snapshot and the digest matches)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("cfl_build", os.path.join(PKG, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()
    yield
