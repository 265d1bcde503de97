from sfmfromscratch_b200.extractor import FeatureExtractor  # noqa: F401
