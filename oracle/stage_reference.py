"""Stage the UNMODIFIED reference into oracle/_ref/ (git-ignored; it travels to the GPU box with gpurun).

TEST / BASELINE INFRASTRUCTURE ONLY.  The reference (reesque/SfmFromScratch) is 14 loose Python files with no
setup.py / pyproject.toml, so `pip install --target baseline/_ref /root/reference` is impossible; this recipe is
the equivalent: it copies the files byte for byte from where they lie under /root/reference into oracle/_ref/
and writes a manifest with their SHA-256 sums.  Nothing under oracle/_ref/ is committed, and nothing in the
product package reads it.  Users:

  * bench.py --impl reference            times the reference's own ScaleRotInvSIFT / NNRatioFeatureMatcher
                                         (cpu_baseline.kind = "reference") beside the oracle port;
  * tests/test_dropin_runner.py          runs the reference's own caller (Runner.FeatureRunner, Runner.py:22-73)
                                         with sfmfromscratch_b200/dropin shadowing the two hot-path packages.

    python oracle/stage_reference.py [--reference /root/reference]
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import shutil

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
FILES = [
    "FeatureExtractor/__init__.py", "FeatureExtractor/FeatureExtractor.py",
    "FeatureExtractor/SIFT/NaiveSIFT.py", "FeatureExtractor/SIFT/ScaleRotInvSIFT.py",
    "FeatureMatcher/__init__.py", "FeatureMatcher/NNRatioFeatureMatcher.py",
    # the hot path's only caller and what it imports (Runner.py:1-18)
    "Runner.py", "SFM.py", "PoseEstimator.py", "Util.py", "Visualizer.py",
]


def stage(reference: str = "/root/reference", force: bool = False) -> str | None:
    """Copy the reference files into oracle/_ref/.  Returns the directory, or None when the reference tree is
    not present (the GPU box: the staged copy that came with the snapshot is used as it is)."""
    if not os.path.isdir(os.path.join(reference, "FeatureExtractor")):
        return REF_DIR if os.path.exists(os.path.join(REF_DIR, "MANIFEST.json")) else None
    manifest = {}
    for rel in FILES:
        src = os.path.join(reference, rel)
        dst = os.path.join(REF_DIR, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if force or not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    with open(os.path.join(REF_DIR, "MANIFEST.json"), "w") as f:
        json.dump({"source": reference, "sha256": manifest}, f, indent=1)
    return REF_DIR


def staged() -> str | None:
    """oracle/_ref/ when a staged reference is there, else None."""
    return REF_DIR if os.path.exists(os.path.join(REF_DIR, "FeatureExtractor", "SIFT", "ScaleRotInvSIFT.py")) else None


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    print(stage(a.reference, a.force))
