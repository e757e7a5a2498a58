"""Training-only RPN label / target assignment (lib/model/rpn/anchor_target_layer.py:31-191) and the RPN
losses built on it (lib/model/rpn/rpn.py:82-105), restated in vectorised torch on the device.
SURVEY.md 8a12: host-side training glue, not a kernel.  Differences from the reference: the random
subsampling uses torch's generator on the tensors' device instead of numpy on the host (same distribution,
no device<->host round trip); everything else -- including the reference's use of the LAST image's example
count for the outside weights (anchor_target_layer.py:153-156) -- is kept."""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from model.utils.config import cfg
from model.utils.net_utils import _smooth_l1_loss
from .bbox_transform import bbox_overlaps_batch, bbox_transform_batch
from .generate_anchors import generate_anchors


def _subsample(mask_row, n_keep, generator=None):
    """indices of `mask_row` to DISABLE so that at most n_keep stay set"""
    idx = torch.nonzero(mask_row).view(-1)
    if idx.numel() <= n_keep:
        return idx[:0]
    perm = torch.randperm(idx.numel(), device=idx.device, generator=generator)
    return idx[perm[: idx.numel() - n_keep]]


class _AnchorTargetLayer(nn.Module):
    def __init__(self, feat_stride, scales, ratios):
        super(_AnchorTargetLayer, self).__init__()
        self._feat_stride = feat_stride
        anchors = torch.from_numpy(generate_anchors(scales=np.array(scales), ratios=np.array(ratios))).float()
        self.register_buffer("_anchors", anchors, persistent=False)
        self._num_anchors = anchors.size(0)
        self._allowed_border = 0
        self.generator = None

    @torch.no_grad()
    def forward(self, input):
        rpn_cls_score, gt_boxes, im_info, num_boxes = input
        B, H, W = gt_boxes.size(0), rpn_cls_score.size(2), rpn_cls_score.size(3)
        dev = gt_boxes.device
        A = self._num_anchors
        sx = torch.arange(W, device=dev, dtype=torch.float32) * self._feat_stride
        sy = torch.arange(H, device=dev, dtype=torch.float32) * self._feat_stride
        yy, xx = torch.meshgrid(sy, sx, indexing="ij")
        shifts = torch.stack((xx.reshape(-1), yy.reshape(-1), xx.reshape(-1), yy.reshape(-1)), 1)
        all_anchors = (self._anchors.to(dev).view(1, A, 4) + shifts.view(-1, 1, 4)).view(-1, 4)
        total = all_anchors.size(0)
        b = self._allowed_border
        keep = ((all_anchors[:, 0] >= -b) & (all_anchors[:, 1] >= -b) & (all_anchors[:, 2] < int(im_info[0][1]) + b) &
                (all_anchors[:, 3] < int(im_info[0][0]) + b))
        inds_inside = torch.nonzero(keep).view(-1)
        anchors = all_anchors[inds_inside]
        n_in = inds_inside.numel()
        labels = gt_boxes.new_full((B, n_in), -1)
        overlaps = bbox_overlaps_batch(anchors, gt_boxes[:, :, :5])
        max_ov, argmax_ov = overlaps.max(2)
        gt_max = overlaps.max(1)[0]
        if not cfg.TRAIN.RPN_CLOBBER_POSITIVES:
            labels[max_ov < cfg.TRAIN.RPN_NEGATIVE_OVERLAP] = 0
        gt_max[gt_max == 0] = 1e-5
        best = overlaps.eq(gt_max.view(B, 1, -1)).sum(2)
        labels[best > 0] = 1
        labels[max_ov >= cfg.TRAIN.RPN_POSITIVE_OVERLAP] = 1
        if cfg.TRAIN.RPN_CLOBBER_POSITIVES:
            labels[max_ov < cfg.TRAIN.RPN_NEGATIVE_OVERLAP] = 0
        num_fg = int(cfg.TRAIN.RPN_FG_FRACTION * cfg.TRAIN.RPN_BATCHSIZE)
        for i in range(B):
            labels[i][_subsample(labels[i] == 1, num_fg, self.generator)] = -1
            n_bg = cfg.TRAIN.RPN_BATCHSIZE - int((labels[i] == 1).sum())
            labels[i][_subsample(labels[i] == 0, n_bg, self.generator)] = -1
        gt_sel = torch.gather(gt_boxes[:, :, :4], 1, argmax_ov.unsqueeze(2).expand(B, n_in, 4))
        bbox_targets = bbox_transform_batch(anchors, gt_sel)
        inside_w = gt_boxes.new_zeros(B, n_in)
        inside_w[labels == 1] = cfg.TRAIN.RPN_BBOX_INSIDE_WEIGHTS[0]
        assert cfg.TRAIN.RPN_POSITIVE_WEIGHT < 0
        num_examples = (labels[B - 1] >= 0).sum().float()          # sic: last image (anchor_target_layer.py:154)
        outside_w = gt_boxes.new_zeros(B, n_in)
        outside_w[labels >= 0] = 1.0 / num_examples

        def unmap(data, fill):
            shape = (B, total) + tuple(data.shape[2:])
            out = data.new_full(shape, fill)
            out[:, inds_inside] = data
            return out

        labels = unmap(labels, -1).view(B, H, W, A).permute(0, 3, 1, 2).contiguous().view(B, 1, A * H, W)
        bbox_targets = unmap(bbox_targets, 0).view(B, H, W, A * 4).permute(0, 3, 1, 2).contiguous()
        iw = unmap(inside_w, 0).view(B, total, 1).expand(B, total, 4).contiguous().view(B, H, W, 4 * A).permute(0, 3, 1, 2).contiguous()
        ow = unmap(outside_w, 0).view(B, total, 1).expand(B, total, 4).contiguous().view(B, H, W, 4 * A).permute(0, 3, 1, 2).contiguous()
        return [labels, bbox_targets, iw, ow]


def rpn_losses(rpn, rpn_cls_score, rpn_bbox_pred, gt_boxes, im_info, num_boxes):
    """rpn.py:82-105: cross-entropy over the sampled anchors + smooth-L1 (sigma 3) on the box deltas."""
    if rpn.RPN_anchor_target is None:
        rpn.RPN_anchor_target = _AnchorTargetLayer(rpn.feat_stride, rpn.anchor_scales, rpn.anchor_ratios)
    B = rpn_cls_score.size(0)
    labels, targets, iw, ow = rpn.RPN_anchor_target((rpn_cls_score.detach(), gt_boxes[:, :, :5], im_info, num_boxes))
    score = rpn.reshape(rpn_cls_score, 2).permute(0, 2, 3, 1).contiguous().view(B, -1, 2)
    label = labels.view(B, -1)
    keep = torch.nonzero(label.view(-1) != -1).view(-1)
    loss_cls = F.cross_entropy(score.view(-1, 2)[keep], label.view(-1)[keep].long())
    loss_box = _smooth_l1_loss(rpn_bbox_pred, targets, iw, ow, sigma=3, dim=[1, 2, 3])
    return loss_cls, loss_box
