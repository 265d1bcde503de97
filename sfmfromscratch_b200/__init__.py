"""sfmfromscratch_b200 -- B200 (sm_100a) implementation of the SfmFromScratch
feature hot path: SIFT-style extraction and NN-ratio matching.

The CUDA library (libsfmb200.so, C ABI in include/sfmb200.h) is the only
compute path; importing the package does not load it, using it does and fails
loudly when the library or a B200 is missing.
"""
from .extractor import FeatureExtractor, NaiveSIFT, ScaleRotInvSIFT, extract_batch, extract_batch_device, make_params
from .matcher import NNRatioFeatureMatcher, match_batch_device, match_device

__all__ = ["FeatureExtractor", "NaiveSIFT", "ScaleRotInvSIFT", "NNRatioFeatureMatcher",
           "extract_batch", "extract_batch_device", "make_params", "match_device", "match_batch_device"]
