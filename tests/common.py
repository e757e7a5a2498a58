"""Seeded synthetic inputs shared by the tests, the golden generators and bench.py
(SURVEY.md section 8d).  numpy only -- no torch, no GPU."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "pytorch-detect-to-track_b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

from d2t_b200.synth import (make_clustered_dets, make_dets, make_frame, make_gt_boxes, make_rois, make_rpn_inputs,  # noqa: E402,F401
                            randn)
from d2t_b200.detect import detect_reference_loop  # noqa: E402,F401
