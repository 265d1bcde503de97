from sfmfromscratch_b200.extractor import NaiveSIFT  # noqa: F401
