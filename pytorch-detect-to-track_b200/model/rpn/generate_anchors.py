"""Anchor table of lib/model/rpn/generate_anchors.py:45-104, restated.

The table is a pure function of (base_size, ratios, scales): for each aspect ratio the 16x16
base box is reshaped to (round(sqrt(area/ratio)), round(w*ratio)) about its centre, then each
is scaled.  Values are 0-based (the comment table at generate_anchors.py:19-37 is 1-based).
"""
import numpy as np


def _centre_form(box):
    w = box[2] - box[0] + 1.0
    h = box[3] - box[1] + 1.0
    return w, h, box[0] + 0.5 * (w - 1.0), box[1] + 0.5 * (h - 1.0)


def _corner_form(ws, hs, xc, yc):
    ws = np.asarray(ws, dtype=np.float64).reshape(-1, 1)
    hs = np.asarray(hs, dtype=np.float64).reshape(-1, 1)
    half_w, half_h = 0.5 * (ws - 1.0), 0.5 * (hs - 1.0)
    return np.hstack((xc - half_w, yc - half_h, xc + half_w, yc + half_h))


def generate_anchors(base_size=16, ratios=[0.5, 1, 2], scales=2 ** np.arange(3, 6)):
    ratios = np.asarray(ratios, dtype=np.float64)
    scales = np.asarray(scales, dtype=np.float64)
    w, h, xc, yc = _centre_form(np.array([0.0, 0.0, base_size - 1.0, base_size - 1.0]))
    ws = np.round(np.sqrt((w * h) / ratios))
    hs = np.round(ws * ratios)
    per_ratio = _corner_form(ws, hs, xc, yc)
    blocks = []
    for row in per_ratio:
        w, h, xc, yc = _centre_form(row)
        blocks.append(_corner_form(w * scales, h * scales, xc, yc))
    return np.vstack(blocks)
