"""Drop-in for lib/model/rpn/rpn.py:17-107 (same submodule / parameter names:
RPN_Conv, RPN_cls_score, RPN_bbox_pred, RPN_proposal, RPN_anchor_target)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from model.utils.config import cfg
from .proposal_layer import _ProposalLayer


class _RPN(nn.Module):
    """Region proposal network head."""

    def __init__(self, din):
        super(_RPN, self).__init__()
        self.din = din
        self.anchor_scales = cfg.ANCHOR_SCALES
        self.anchor_ratios = cfg.ANCHOR_RATIOS
        self.feat_stride = cfg.FEAT_STRIDE[0]
        n_anchor = len(self.anchor_scales) * len(self.anchor_ratios)
        self.RPN_Conv = nn.Conv2d(self.din, 512, 3, 1, 1, bias=True)
        self.nc_score_out = n_anchor * 2
        self.RPN_cls_score = nn.Conv2d(512, self.nc_score_out, 1, 1, 0)
        self.nc_bbox_out = n_anchor * 4
        self.RPN_bbox_pred = nn.Conv2d(512, self.nc_bbox_out, 1, 1, 0)
        self.RPN_proposal = _ProposalLayer(self.feat_stride, self.anchor_scales, self.anchor_ratios)
        self.RPN_anchor_target = None   # built lazily by model.rpn.anchor_target_layer in training
        self.rpn_loss_cls = 0
        self.rpn_loss_box = 0

    @staticmethod
    def reshape(x, d):
        s = x.size()
        return x.view(s[0], int(d), int(float(s[1] * s[2]) / float(d)), s[3])

    @staticmethod
    def cls_prob_from_score(rpn_cls_score, nc_score_out):
        """rpn.py:66-68: softmax over {bg, fg} pairs (channel a vs channel A + a)."""
        r = _RPN.reshape(rpn_cls_score, 2)
        return _RPN.reshape(F.softmax(r, dim=1), nc_score_out)

    def forward(self, base_feat, im_info, gt_boxes=None, num_boxes=None):
        rpn_conv1 = F.relu(self.RPN_Conv(base_feat), inplace=True)
        return self.forward_from_maps(self.RPN_cls_score(rpn_conv1), self.RPN_bbox_pred(rpn_conv1), im_info, gt_boxes,
                                      num_boxes)

    def proposals_from_maps(self, rpn_cls_score, rpn_bbox_pred, im_info):
        """rpn.py:66-78: the proposal step alone, for any number of images (they are independent: the training heads run
        it ONCE for both legs of all pairs instead of once per leg -- one NMS launch of 2B lists instead of two of B)"""
        rpn_cls_prob = self.cls_prob_from_score(rpn_cls_score, self.nc_score_out)
        cfg_key = 'TRAIN' if self.training else 'TEST'
        return self.RPN_proposal((rpn_cls_prob.detach(), rpn_bbox_pred.detach(), im_info, cfg_key))

    def forward_from_maps(self, rpn_cls_score, rpn_bbox_pred, im_info, gt_boxes=None, num_boxes=None, rois=None):
        """rpn.py:66-105 after the three convolutions (the training engine runs those on the tcgen05 kernel and hands
        their outputs in as autograd leaves).  `rois`: proposals already computed by proposals_from_maps."""
        if rois is None:
            rois = self.proposals_from_maps(rpn_cls_score, rpn_bbox_pred, im_info)
        self.rpn_loss_cls = 0
        self.rpn_loss_box = 0
        if self.training:
            from .anchor_target_layer import rpn_losses   # training-only glue
            self.rpn_loss_cls, self.rpn_loss_box = rpn_losses(self, rpn_cls_score, rpn_bbox_pred, gt_boxes, im_info,
                                                              num_boxes)
        return rois, self.rpn_loss_cls, self.rpn_loss_box
