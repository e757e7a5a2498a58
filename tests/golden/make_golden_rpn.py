"""Generates tests/golden/rpn_*.npz by IMPORTING THE REFERENCE'S OWN PYTHON (CPU, this container
only: /root/reference does not exist on the GPU box).  Recipe: SURVEY.md App. D.3 -- put
/root/reference/lib on sys.path and shim easydict, model.nms.nms_gpu (-> the C NMS oracle, itself
pinned against the reference CUDA kernel by nms_*.npz) and model.roi_crop.functions.roi_crop.

    python tests/golden/make_golden_rpn.py
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import common  # noqa: E402  (adds the repo root to sys.path; we then hide OUR `model` package)

sys.path = [p for p in sys.path if p != common.PKG]
sys.path.insert(0, "/root/reference/lib")
from oracle import cpu as oracle  # noqa: E402


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
m = types.ModuleType("model.nms.nms_gpu")
m.nms_gpu = lambda dets, thresh: torch.from_numpy(oracle.nms(dets.numpy(), float(thresh))).view(-1, 1)
sys.modules["model.nms.nms_gpu"] = m
m = types.ModuleType("model.roi_crop.functions.roi_crop"); m.RoICropFunction = object
sys.modules["model.roi_crop.functions.roi_crop"] = m

from model.utils.config import cfg  # noqa: E402
from model.rpn.generate_anchors import generate_anchors  # noqa: E402
from model.rpn.proposal_layer import _ProposalLayer  # noqa: E402
from model.rpn.bbox_transform import bbox_transform_inv, clip_boxes  # noqa: E402

out = {}
out["anchors_default"] = generate_anchors()
out["anchors_d2t"] = generate_anchors(scales=np.array([4, 8, 16, 32]), ratios=np.array([0.5, 1, 2]))

cfg.ANCHOR_SCALES = [4, 8, 16, 32]
layer = _ProposalLayer(16, cfg.ANCHOR_SCALES, cfg.ANCHOR_RATIOS)
cases = {"small": dict(B=2, H=10, W=14, seed=30), "config1": dict(B=1, H=19, W=32, seed=31, im_h=300, im_w=500),
         "full": dict(B=1, H=38, W=63, seed=32, im_h=600, im_w=1000)}
for name, kw in cases.items():
    prob, deltas, im_info = common.make_rpn_inputs(**kw)
    for key in ("TEST", "TRAIN"):
        rois = layer((torch.from_numpy(prob), torch.from_numpy(deltas), torch.from_numpy(im_info), key))
        out["rois_%s_%s" % (name, key)] = rois.numpy()
# decode + clip on their own (bbox_transform.py:108-134, 156-173)
rng = np.random.RandomState(33)
boxes = torch.from_numpy(common.make_rois(64, 1, seed=34)[None, :, 1:].copy())
deltas = torch.from_numpy((rng.standard_normal((1, 64, 4)) * 0.5).astype(np.float32))
pred = bbox_transform_inv(boxes, deltas, 1)
out["bti_boxes"], out["bti_deltas"], out["bti_pred"] = boxes.numpy(), deltas.numpy(), pred.numpy().copy()
out["bti_clipped"] = clip_boxes(pred.clone(), torch.tensor([[600., 1000., 1.]]), 1).numpy()
np.savez_compressed(os.path.join(HERE, "rpn_reference.npz"), **out)
print({k: v.shape for k, v in out.items()})
