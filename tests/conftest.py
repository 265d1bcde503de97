import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = os.environ.get("SFM_REFERENCE", "/root/reference")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _cuda_ok():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # gpu tests never run (and never silently pass) on a box without CUDA
    if _cuda_ok():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def reference_modules():
    """The unmodified reference, imported from /root/reference (build container only)."""
    if not os.path.isdir(os.path.join(REFERENCE, "FeatureExtractor")):
        pytest.skip("reference tree not present (GPU box)")
    import importlib
    saved = list(sys.path)
    saved_mods = {k: v for k, v in sys.modules.items() if k.split(".")[0] in ("FeatureExtractor", "FeatureMatcher")}
    for k in saved_mods:
        del sys.modules[k]
    sys.path.insert(0, REFERENCE)
    try:
        fe = importlib.import_module("FeatureExtractor")
        fm = importlib.import_module("FeatureMatcher")
        assert fe.__file__.startswith(REFERENCE)
        yield fe, fm
    finally:
        sys.path[:] = saved
        for k in [k for k in sys.modules if k.split(".")[0] in ("FeatureExtractor", "FeatureMatcher")]:
            del sys.modules[k]
        sys.modules.update(saved_mods)
