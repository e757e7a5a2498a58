"""Generates tests/golden/targets_reference.npz by IMPORTING THE REFERENCE'S OWN PYTHON target layers
(CPU, this container only).  Shims as in make_golden_rpn.py.   python tests/golden/make_golden_targets.py"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import common  # noqa: E402

sys.path = [p for p in sys.path if p != common.PKG]
sys.path.insert(0, "/root/reference/lib")
import builtins  # noqa: E402
builtins.long = int            # python-2 name used by anchor_target_layer.py:83-84


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
m = types.ModuleType("model.nms.nms_gpu"); m.nms_gpu = None; sys.modules["model.nms.nms_gpu"] = m
m = types.ModuleType("model.roi_crop.functions.roi_crop"); m.RoICropFunction = object
sys.modules["model.roi_crop.functions.roi_crop"] = m

from model.utils.config import cfg  # noqa: E402
from model.rpn.anchor_target_layer import _AnchorTargetLayer  # noqa: E402
from model.rpn.tracking_proposal_target_layer import _TrackingProposalTargetLayer  # noqa: E402
from model.rpn.bbox_transform import bbox_overlaps_batch, bbox_transform_batch  # noqa: E402

cfg.ANCHOR_SCALES = [4, 8, 16, 32]
cfg.TRAIN.BATCH_SIZE = 128
out = {}
gt = common.make_gt_boxes(B=2, K=30, seed=2, height=300, width=500)        # [B, 2, K, 6]
out["gt"] = gt
# ---- anchor target layer, leg 0
np.random.seed(0)
layer = _AnchorTargetLayer(16, cfg.ANCHOR_SCALES, cfg.ANCHOR_RATIOS)
score = torch.zeros(2, 24, 19, 32)
g = torch.from_numpy(gt[:, 0, :, :5].copy())
nb = torch.from_numpy((gt[:, 0, :, 4] > 0).sum(1).astype(np.int64)).view(2, 1)
info = torch.tensor([[300., 500., 1.], [300., 500., 1.]])
lab, tgt, iw, ow = layer((score, g, info, nb))
out["at_labels"], out["at_targets"], out["at_iw"], out["at_ow"] = lab.numpy(), tgt.numpy(), iw.numpy(), ow.numpy()
# ---- box helpers
rois = torch.from_numpy(common.make_rois(50, 1, 300, 500, seed=5)[:, 1:].copy())
out["ov_anchors2d"] = bbox_overlaps_batch(rois, g).numpy()
rois3 = torch.from_numpy(np.stack([common.make_rois(40, 1, 300, 500, seed=6 + i) for i in range(2)]))
out["ov_rois3d"] = bbox_overlaps_batch(rois3, g).numpy()
ex = rois3[:, :30, 1:5].contiguous()
out["bt_batch"] = bbox_transform_batch(ex, g[:, :, :4].contiguous() + 1.0).numpy()
# ---- tracking targets (deterministic)
tl = _TrackingProposalTargetLayer(31)
gt_l = torch.from_numpy(gt.transpose(1, 0, 2, 3).copy())
nb_l = torch.from_numpy((gt.transpose(1, 0, 2, 3)[..., 4] > 0).sum(-1).astype(np.int64)).view(2, 2, 1)
r, l, t, iw, ow = tl(gt_l, nb_l)
out["trk_rois"], out["trk_labels"], out["trk_targets"], out["trk_iw"], out["trk_ow"] = (x.numpy() for x in (r, l, t, iw, ow))
np.savez_compressed(os.path.join(HERE, "targets_reference.npz"), **out)
print({k: v.shape for k, v in out.items()})
