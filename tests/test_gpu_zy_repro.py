"""Network-level calibration is bit-reproducible: the BraTS miniature calibrated twice in one process gives identical
bit patterns for every intermediate of every layer (input, activation scale, codes, A0, B0, the five inverses, the
200-iterate loss history, best weights, layer output).  Same code as tools/repro_check.py, whose round-1 run is
profiles/r01_repro_check.txt."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_calibrations_of_the_miniature_are_bit_identical():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import repro_check
    rec1, loss1 = repro_check.run_once("brats")
    rec2, loss2 = repro_check.run_once("brats")
    assert len(rec1) == len(rec2) and len(rec1) > 100
    assert repro_check.compare(rec1, rec2, "brats miniature, run 1 vs run 2")
    assert loss1 == loss2
