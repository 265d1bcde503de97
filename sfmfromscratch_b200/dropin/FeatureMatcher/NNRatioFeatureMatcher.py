from sfmfromscratch_b200.matcher import NNRatioFeatureMatcher  # noqa: F401
