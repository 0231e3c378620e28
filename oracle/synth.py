"""Moved: the synthetic input generators live in hgmm_b200/synth.py (they are bench / test INPUT generators, not part of
the oracle).  This alias keeps `from oracle import synth` working for the tests."""
import os
import sys

_PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                    "gpu-accelerated-point-cloud-registration-using-hierarchical-gmm_b200")
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)
from hgmm_b200.synth import bunny_like, lidar_sweep  # noqa: E402,F401
