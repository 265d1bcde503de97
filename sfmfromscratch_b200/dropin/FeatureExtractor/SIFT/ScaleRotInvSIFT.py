"""main.py:10 imports `FeatureExtractor.SIFT.ScaleRotInvSIFT.ScaleRotInvSIFT`."""
from sfmfromscratch_b200.extractor import ScaleRotInvSIFT  # noqa: F401
