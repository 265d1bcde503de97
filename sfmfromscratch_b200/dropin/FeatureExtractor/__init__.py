"""Drop-in for the reference's top-level `FeatureExtractor` package
(FeatureExtractor/__init__.py:1-3): put `sfmfromscratch_b200/dropin` ahead of the
reference on sys.path and Runner.py / main.py import these classes unchanged."""
from sfmfromscratch_b200.extractor import FeatureExtractor, NaiveSIFT, ScaleRotInvSIFT  # noqa: F401
