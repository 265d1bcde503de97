"""Drop-in for the reference's `FeatureMatcher` package (FeatureMatcher/__init__.py:1)."""
from sfmfromscratch_b200.matcher import NNRatioFeatureMatcher  # noqa: F401
