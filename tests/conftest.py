"""pytest configuration.

Markers: ``gpu`` = needs a real B200 (run with ``-m gpu`` through gpurun); everything else
runs on the CPU-only container.  The product package directory is named
``ssnt-tts-rust_b200`` (not an importable identifier), so it is loaded here under the module
name ``ssnt_tts_rust_b200``.
"""
import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def load_product():
    """Import ``ssnt-tts-rust_b200/`` as module ``ssnt_tts_rust_b200``."""
    name = "ssnt_tts_rust_b200"
    if name in sys.modules:
        return sys.modules[name]
    pkg_dir = os.path.join(ROOT, "ssnt-tts-rust_b200")
    spec = importlib.util.spec_from_file_location(
        name, os.path.join(pkg_dir, "__init__.py"), submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (B200); run via gpurun")
    config.addinivalue_line("markers", "timeout: per-test limit (pytest-timeout) for kernels that could hang")


def has_cuda() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def product():
    return load_product()


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


def pytest_collection_modifyitems(config, items):
    if has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (run with gpurun)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
