"""Seeded inputs of the golden fixtures -- shared by make_golden_gpu.py (which runs the
reference's own kernels on them) and tests/test_oracle_golden.py (which runs the CPU oracle)."""
import numpy as np

import common


def psroi_cases():
    H, W, D = 20, 30, 4
    rois = common.make_rois(48, 2, height=H * 16, width=W * 16, seed=40, lo=8.0, hi=400.0, shuffle=True)
    # degenerate / boundary rois: zero-size, inverted, far outside, .5 coordinates (roundf half-away)
    extra = np.array([[0, 10.5, 10.5, 10.5, 10.5], [1, 100.5, 50.5, 20.5, 30.5], [0, 600, 400, 700, 500],
                      [1, 0, 0, 479, 319], [0, 31.5, 47.5, 63.5, 95.5], [1, -40, -40, 10, 10]], np.float32)
    rois = np.concatenate([rois, extra], 0)
    feat = common.randn((2, D * 49, H, W), 41)
    gtop = common.randn((rois.shape[0], D, 7, 7), 42)
    return {"r7": dict(feat=feat, rois=rois, gtop=gtop, scale=1.0 / 16.0, P=7, G=7, D=D)}


def nms_cases():
    return {
        "n300_t03": (common.make_dets(300, seed=50), 0.3),
        "n2000_t07": (common.make_dets(2000, seed=51), 0.7),
        "clustered1000_t07": (common.make_clustered_dets(1000, seed=52), 0.7),
        "clustered130_t05": (common.make_clustered_dets(130, seed=53), 0.5),
        "n1": (common.make_dets(1, seed=54), 0.7),
        "n65": (common.make_dets(65, seed=55, height=120, width=160), 0.3),
    }


def corr_cases():
    out = {}
    for name, (C, H, W, p) in {"d2t_s1": (32, 10, 14, (8, 1, 8, 1, 1)), "d2t_s2": (16, 21, 27, (8, 1, 8, 2, 2)),
                               "k3": (8, 12, 13, (4, 3, 4, 1, 2)), "pad0": (8, 24, 26, (0, 1, 4, 1, 1))}.items():
        seed = 60 + len(out)
        a, b = common.randn((2, C, H, W), seed), common.randn((2, C, H, W), seed + 100)
        out[name] = dict(in1=a, in2=b, params=p, gseed=seed + 200)
    return out


def roi_cases():
    H, W, C = 20, 30, 6
    rois = common.make_rois(40, 2, height=H * 16, width=W * 16, seed=70, lo=8.0, hi=400.0)
    rois = np.concatenate([rois, np.array([[0, -30, -30, 5, 5], [1, 470, 310, 500, 340]], np.float32)], 0)
    feat = common.randn((2, C, H, W), 71)
    rng = np.random.RandomState(72)
    R = 2 * 5
    theta = np.concatenate([rng.uniform(0.2, 0.8, (R, 1)), np.zeros((R, 1)), rng.uniform(-0.5, 0.5, (R, 1)),
                            np.zeros((R, 1)), rng.uniform(0.2, 0.8, (R, 1)), rng.uniform(-0.5, 0.5, (R, 1))], 1)
    lin = np.linspace(-1, 1, 7)
    gx, gy = np.meshgrid(lin, lin)
    grid = np.zeros((R, 7, 7, 2), np.float32)
    for r in range(R):
        t = theta[r].reshape(2, 3)
        x = t[0, 0] * gx + t[0, 1] * gy + t[0, 2]
        y = t[1, 0] * gx + t[1, 1] * gy + t[1, 2]
        grid[r, :, :, 0], grid[r, :, :, 1] = y, x          # (y, x) order, faster_rcnn.py:75-77
    grid[0] *= 2.5                                          # partly outside [-1, 1]
    return dict(feat=feat, rois=rois, grid=grid, scale=1.0 / 16.0)
