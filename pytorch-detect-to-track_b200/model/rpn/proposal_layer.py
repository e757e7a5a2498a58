"""Drop-in for lib/model/rpn/proposal_layer.py:27-161.

Same constructor and call shape -- ``_ProposalLayer(feat_stride, scales, ratios)((rpn_cls_prob,
rpn_bbox_pred, im_info, cfg_key))`` -> rois ``[B, post_nms_topN, 5]`` -- but the whole batch goes
through four kernel launches + one sort (d2t_b200.ops.proposals) instead of a numpy shift table,
~20 elementwise torch kernels and a python loop with a device<->host NMS round trip per image.
"""
import numpy as np
import torch
import torch.nn as nn

from d2t_b200 import ops
from model.utils.config import cfg
from .generate_anchors import generate_anchors


class _ProposalLayer(nn.Module):
    def __init__(self, feat_stride, scales, ratios):
        super(_ProposalLayer, self).__init__()
        self._feat_stride = feat_stride
        anchors = torch.from_numpy(generate_anchors(scales=np.array(scales), ratios=np.array(ratios))).float()
        self.register_buffer("_anchors", anchors, persistent=False)
        self._num_anchors = anchors.size(0)

    def forward(self, input):
        cls_prob, bbox_deltas, im_info, cfg_key = input
        pre_nms_topN = cfg[cfg_key].RPN_PRE_NMS_TOP_N
        post_nms_topN = cfg[cfg_key].RPN_POST_NMS_TOP_N
        nms_thresh = cfg[cfg_key].RPN_NMS_THRESH
        anchors = self._anchors
        if anchors.device != bbox_deltas.device:
            anchors = anchors.to(bbox_deltas.device)
        with torch.no_grad():
            return ops.proposals(anchors, bbox_deltas.contiguous(), cls_prob.contiguous(),
                                 im_info.contiguous().float(), self._feat_stride, pre_nms_topN, post_nms_topN,
                                 nms_thresh)
