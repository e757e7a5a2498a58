"""Generates tests/golden/minibatch_reference.npz by running THE REFERENCE'S OWN `get_minibatch`
(lib/roi_data_layer/minibatch.py:20-56 -> _get_image_blob :58-88 -> lib/model/utils/blob.py) on two synthetic frames
written as lossless PNGs -- this container only: /root/reference does not exist on the GPU box.  cv2 runs with IPP off
(OpenCV's own resize code, see make_golden_frames.py); `scipy.misc.imread` (gone from scipy) is shimmed, it is imported but
never called on this path.

    python tests/golden/make_golden_minibatch.py
"""
import os
import sys
import tempfile
import types

import cv2
import numpy as np
import scipy.sparse

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import common  # noqa: E402

sys.path = [p for p in sys.path if p != common.PKG]
for k in [k for k in sys.modules if k == "model" or k.startswith("model.")]:
    del sys.modules[k]                               # OUR model package (imported by common): hide it
sys.path.insert(0, "/root/reference/lib")


class EasyDict(dict):
    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, EasyDict):
            v = EasyDict(v)
        super().__setitem__(k, v)

    __setattr__ = __setitem__

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


m = types.ModuleType("easydict"); m.EasyDict = EasyDict; sys.modules["easydict"] = m
import scipy.misc  # noqa: E402
scipy.misc.imread = None
np.object = object                                   # (aliases the 2018 code uses)

from model.utils.config import cfg  # noqa: E402
from roi_data_layer.minibatch import get_minibatch  # noqa: E402

cv2.ipp.setUseIPP(False)
cfg.TRAIN.SCALES = (60,)                             # small fixtures; the rule (shorter side -> SCALES[0]) is the same
out = {}
with tempfile.TemporaryDirectory() as tmp:
    for name, (h, w, seed, flipped) in {"landscape": (45, 80, 81, False), "flipped": (45, 80, 82, True),
                                        "portrait": (70, 40, 83, False)}.items():
        im = common.make_frame(h, w, seed)
        path = os.path.join(tmp, name + ".png")
        cv2.imwrite(path, im)
        rng = np.random.RandomState(seed)
        n = 4
        x1, y1 = rng.randint(0, w // 2, n), rng.randint(0, h // 2, n)
        boxes = np.stack([x1, y1, x1 + rng.randint(4, w // 2, n), y1 + rng.randint(4, h // 2, n)], 1).astype(np.uint16)
        classes = np.array([3, 0, 17, 30], np.int32)  # one background row: dropped (minibatch.py:39-41)
        overlaps = np.zeros((n, 31), np.float32)
        overlaps[np.arange(n), classes] = 1.0
        roidb = [{"image": path, "flipped": flipped, "boxes": boxes, "gt_classes": classes, "track_id": np.array([5, 6, 7, 8]),
                  "gt_overlaps": scipy.sparse.csr_matrix(overlaps), "img_id": 1000 + seed, "width": w, "height": h}]
        blobs = get_minibatch(roidb, 31)
        out[name + "_im"], out[name + "_boxes"], out[name + "_classes"] = im, boxes, classes
        out[name + "_track_id"], out[name + "_flipped"] = roidb[0]["track_id"], np.array(flipped)
        out[name + "_data"], out[name + "_gt_boxes"] = blobs["data"], blobs["gt_boxes"]
        out[name + "_im_info"], out[name + "_img_id"] = blobs["im_info"], np.array(blobs["img_id"])
        print(name, blobs["data"].shape, blobs["gt_boxes"].shape, blobs["im_info"])
np.savez_compressed(os.path.join(HERE, "minibatch_reference.npz"), **out)
