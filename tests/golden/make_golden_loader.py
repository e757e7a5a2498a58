"""Generates tests/golden/loader_reference.npz by running THE REFERENCE'S OWN `roibatchLoader.__getitem__`
(lib/roi_data_layer/roibatchLoader.py:25-269, imported from its source with the tabs expanded -- it mixes tabs and spaces)
on synthetic frame pairs written as lossless PNGs; `.cuda()` is a no-op here (the reference moves num_boxes to the GPU).
This container only.  cv2 with IPP off (see make_golden_frames.py); numpy's global generator seeded per sample.

    python tests/golden/make_golden_loader.py
"""
import os
import sys
import tempfile
import types

import cv2
import numpy as np
import scipy.sparse
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import common  # noqa: E402

sys.path = [p for p in sys.path if p != common.PKG]
for k in [k for k in sys.modules if k == "model" or k.startswith("model.") or k.startswith("roi_data_layer")]:
    del sys.modules[k]
sys.path.insert(0, "/root/reference/lib")
from make_golden_minibatch_shims import install  # noqa: E402
install()
torch.Tensor.cuda = lambda self, *a, **k: self

from model.utils.config import cfg  # noqa: E402
src = open("/root/reference/lib/roi_data_layer/roibatchLoader.py").read().expandtabs(8)
mod = types.ModuleType("ref_roibatchLoader")
exec(compile(src, "roibatchLoader.py", "exec"), mod.__dict__)

cv2.ipp.setUseIPP(False)
cfg.TRAIN.SCALES = (60,)
cfg.MAX_NUM_GT_BOXES = 30                               # trainval_net.py:165


def entry(tmp, name, h, w, seed, flipped, need_crop, n=4, tiny=False):
    im = common.make_frame(h, w, seed)
    path = os.path.join(tmp, "%s_%d.png" % (name, seed))
    cv2.imwrite(path, im)
    rng = np.random.RandomState(seed)
    x1, y1 = rng.randint(2, w // 2, n), rng.randint(2, h // 2, n)
    boxes = np.stack([x1, y1, x1 + rng.randint(4, w // 2 - 2, n), y1 + rng.randint(4, h // 2 - 2, n)], 1).astype(np.uint16)
    if tiny:
        boxes[1] = [w - 2, 3, w - 1, 9]                  # falls outside a width crop: degenerate after clamping
    classes = np.array([3, 17, 30, 5][:n], np.int32)
    ov = np.zeros((n, 31), np.float32)
    ov[np.arange(n), classes] = 1.0
    return {"image": path, "flipped": flipped, "boxes": boxes, "gt_classes": classes, "track_id": np.arange(n) + 1,
            "gt_overlaps": scipy.sparse.csr_matrix(ov), "img_id": seed, "width": w, "height": h, "need_crop": need_crop}, im


# name -> (h, w, ratio given to the loader, need_crop, training, tiny)
CASES = {"pad_landscape": (45, 80, 2.0, 0, True, False), "pad_portrait": (70, 40, 0.5, 0, True, False),
         "square": (45, 80, 1.0, 0, True, False), "crop_wide": (30, 120, 2.0, 1, True, True),
         "crop_tall": (120, 30, 0.5, 1, True, False), "eval": (45, 80, 1.7777, 0, False, False)}
out = {}
with tempfile.TemporaryDirectory() as tmp:
    for ci, (name, (h, w, ratio, need_crop, training, tiny)) in enumerate(CASES.items()):
        e0, im0 = entry(tmp, name, h, w, 200 + 2 * ci, False, need_crop, tiny=tiny)
        e1, im1 = entry(tmp, name, h, w, 201 + 2 * ci, False, need_crop, tiny=tiny)
        loader = mod.roibatchLoader([[e0, e1]], [ratio], [0], 1, 31, training=training)
        np.random.seed(500 + ci)
        data, im_info, gt, num = loader[0]
        out[name + "_im0"], out[name + "_im1"] = im0, im1
        for k in ("boxes", "gt_classes", "track_id"):
            out["%s_%s0" % (name, k)], out["%s_%s1" % (name, k)] = e0[k], e1[k]
        out[name + "_cfg"] = np.array([ratio, need_crop, int(training), 500 + ci, 200 + 2 * ci])
        out[name + "_data"], out[name + "_im_info"] = data.numpy(), im_info.numpy()
        out[name + "_gt"], out[name + "_num"] = gt.numpy(), num.numpy()
        print(name, tuple(data.shape), im_info.tolist(), num.view(-1).tolist())
np.savez_compressed(os.path.join(HERE, "loader_reference.npz"), **out)
