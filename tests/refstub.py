"""Helpers for running the reference's own caller (Runner.py) in tests: a stub `matplotlib` (not installed in this
image; Runner.py:5 and Visualizer.py:1-4 import it for plotting only) and a context manager that puts a chosen
set of directories in front of sys.path while purging the reference's top-level module names, so the same
`import Runner` can be resolved against the unmodified reference or against sfmfromscratch_b200/dropin."""
import contextlib
import sys
import types

_REF_NAMES = ("FeatureExtractor", "FeatureMatcher", "Runner", "SFM", "PoseEstimator", "Util", "Visualizer")


def install_matplotlib_stub():
    if "matplotlib" in sys.modules and not getattr(sys.modules["matplotlib"], "_sfm_stub", False):
        return
    mpl = types.ModuleType("matplotlib")
    mpl._sfm_stub = True
    for sub in ("pyplot", "cm", "widgets"):
        m = types.ModuleType("matplotlib." + sub)
        setattr(mpl, sub, m)
        sys.modules["matplotlib." + sub] = m
    sys.modules["matplotlib.widgets"].Button = type("Button", (), {})
    sys.modules["matplotlib"] = mpl


@contextlib.contextmanager
def reference_path(*dirs):
    """sys.path = dirs + old path, with the reference's module names purged before and after."""
    saved_path = list(sys.path)
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in _REF_NAMES}
    for k in saved:
        del sys.modules[k]
    install_matplotlib_stub()
    sys.path[:0] = list(dirs)
    try:
        yield
    finally:
        sys.path[:] = saved_path
        for k in [k for k in sys.modules if k.split(".")[0] in _REF_NAMES]:
            del sys.modules[k]
        sys.modules.update(saved)
