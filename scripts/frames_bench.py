"""Device time of d2t_frames_prep for the step's four 720x1280 frames (cold L2), both layouts."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import common  # noqa: E402
from d2t_b200 import ops  # noqa: E402

flush = torch.zeros(128 * 1024 * 1024, device="cuda")
fr = torch.from_numpy(np.stack([common.make_frame(720, 1280, 60 + i) for i in range(4)])).cuda()
for cap in (True, False):
    fh, fw, fs = ops.frames_resized_shape(720, 1280, 600, 1000, cap)
    for nhwc in (False, True):
        out = torch.empty((4, fh, fw, 3) if nhwc else (4, 3, fh, fw), device="cuda")
        ms = bench.time_kernel(lambda: ops.frames_prep(fr, fs, out=out, nhwc=nhwc), 20, flush)
        nbytes = fr.numel() + 4 * out.numel()
        print("frames_prep 4x720x1280 -> %dx%d nhwc=%d: %.2f us  %.0f GB/s" % (fh, fw, nhwc, ms * 1e3, nbytes / ms / 1e6))
