"""Development aid: the GPU test suite with SFM_OPT_HARRIS_STREAM_MIN_BANDS = 0, i.e. every pyramid level the
persistent Harris stream can take (W % 4 == 0, W >= 76, 7x7 window) goes through it instead of the tile kernel."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sfmfromscratch_b200 import _native as N

N.load_library()
N.set_option(N.SFM_OPT_HARRIS_STREAM_MIN_BANDS, 0)
sys.exit(pytest.main(["-m", "gpu", "-x", "-q", os.path.join(ROOT, "tests"), "--deselect", "tests/test_gpu_harris_stream.py"]))
