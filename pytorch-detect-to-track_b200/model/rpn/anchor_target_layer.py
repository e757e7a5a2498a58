"""Training-only RPN label / target assignment (lib/model/rpn/anchor_target_layer.py:31-191) and the RPN
losses built on it (lib/model/rpn/rpn.py:82-105), restated in vectorised torch on the device.
SURVEY.md 8a12: host-side training glue, not a kernel.

Written WITHOUT any device->host round trip and with static shapes only (no ``nonzero``, no ``int(tensor)``), so that the
training step's head -- target layers, losses and their autograd backward -- can be replayed as one CUDA graph
(d2t_b200.train).  Consequences, none of which changes a result:
  * anchors outside the image are not compacted away; they carry label -1 / zero targets from the start and are masked
    out of the per-ground-truth maxima (the reference computes on the inside subset and un-maps at the end);
  * "keep a random subset of n" is a sort of uniform random keys over the candidates and a threshold at the n-th key
    (the reference permutes the candidate indices with numpy on the host): the same distribution, no host RNG.
Kept from the reference: the negatives' quota is RPN_BATCHSIZE minus the number of positives BEFORE their subsampling
(anchor_target_layer.py:119,137), and the outside weights use the LAST image's example count (:153-156)."""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from model.utils.config import cfg
from model.utils.net_utils import _smooth_l1_loss
from .bbox_transform import bbox_overlaps_batch, bbox_transform_batch
from .generate_anchors import generate_anchors


def _keep_random(mask, n_keep, generator=None):
    """mask [B, N] bool, n_keep: int or long tensor [B] -> the candidates that STAY: a uniformly random subset of
    min(n_keep, count) of them per row.  No host synchronisation."""
    B, N = mask.shape
    keys = torch.rand(B, N, device=mask.device, generator=generator)
    keys = torch.where(mask, keys, torch.full_like(keys, 2.0))            # non-candidates sort last
    srt = torch.sort(keys, dim=1)[0]
    if not torch.is_tensor(n_keep):
        n_keep = torch.full((B,), int(n_keep), device=mask.device, dtype=torch.long)
    n_keep = n_keep.clamp(min=0, max=N)
    thr = srt.gather(1, (n_keep - 1).clamp(min=0).view(B, 1))             # the n-th smallest key (2.0 if fewer candidates)
    return mask & (keys <= thr) & (n_keep > 0).view(B, 1)


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
        if self._anchors.device != dev:
            self._anchors = self._anchors.to(dev)              # once (the layer is created lazily, on the host)
        anchors = (self._anchors.view(1, A, 4) + shifts.view(-1, 1, 4)).view(-1, 4)
        total = anchors.size(0)
        b = self._allowed_border
        im_w, im_h = torch.trunc(im_info[0][1]), torch.trunc(im_info[0][0])           # long(im_info[0][.]) in the reference
        inside = ((anchors[:, 0] >= -b) & (anchors[:, 1] >= -b) & (anchors[:, 2] < im_w + b) & (anchors[:, 3] < im_h + b))
        ins = inside.view(1, total)
        overlaps = bbox_overlaps_batch(anchors, gt_boxes[:, :, :5])                   # [B, total, K]
        max_ov, argmax_ov = overlaps.max(2)
        gt_max = overlaps.masked_fill(~inside.view(1, total, 1), -1.0).max(1)[0]      # per ground truth, inside anchors only
        labels = gt_boxes.new_full((B, total), -1)
        if not cfg.TRAIN.RPN_CLOBBER_POSITIVES:
            labels[max_ov < cfg.TRAIN.RPN_NEGATIVE_OVERLAP] = 0
        gt_max = torch.where(gt_max == 0, torch.full_like(gt_max, 1e-5), gt_max)
        best = overlaps.eq(gt_max.view(B, 1, -1)).sum(2)
        labels[best > 0] = 1
        labels[max_ov >= cfg.TRAIN.RPN_POSITIVE_OVERLAP] = 1
        if cfg.TRAIN.RPN_CLOBBER_POSITIVES:
            labels[max_ov < cfg.TRAIN.RPN_NEGATIVE_OVERLAP] = 0
        labels = torch.where(ins, labels, torch.full_like(labels, -1))
        num_fg = int(cfg.TRAIN.RPN_FG_FRACTION * cfg.TRAIN.RPN_BATCHSIZE)
        fg, bg = labels == 1, labels == 0
        sum_fg = fg.sum(1)                                                            # BEFORE the subsampling (see above)
        keep_fg = _keep_random(fg, num_fg, self.generator)
        keep_bg = _keep_random(bg, cfg.TRAIN.RPN_BATCHSIZE - sum_fg, self.generator)
        labels = torch.where((fg & ~keep_fg) | (bg & ~keep_bg), torch.full_like(labels, -1), labels)
        gt_sel = torch.gather(gt_boxes[:, :, :4], 1, argmax_ov.unsqueeze(2).expand(B, total, 4))
        bbox_targets = bbox_transform_batch(anchors, gt_sel) * ins.view(1, total, 1).float()
        inside_w = (labels == 1).float() * cfg.TRAIN.RPN_BBOX_INSIDE_WEIGHTS[0]
        assert cfg.TRAIN.RPN_POSITIVE_WEIGHT < 0
        num_examples = (labels[B - 1] >= 0).sum().float()          # sic: last image (anchor_target_layer.py:154)
        outside_w = (labels >= 0).float() / num_examples
        labels = labels.view(B, H, W, A).permute(0, 3, 1, 2).contiguous().view(B, 1, A * H, W)
        bbox_targets = bbox_targets.view(B, H, W, A * 4).permute(0, 3, 1, 2).contiguous()
        iw = inside_w.view(B, total, 1).expand(B, total, 4).contiguous().view(B, H, W, 4 * A).permute(0, 3, 1, 2).contiguous()
        ow = outside_w.view(B, total, 1).expand(B, total, 4).contiguous().view(B, H, W, 4 * A).permute(0, 3, 1, 2).contiguous()
        return [labels, bbox_targets, iw, ow]


def rpn_losses(rpn, rpn_cls_score, rpn_bbox_pred, gt_boxes, im_info, num_boxes):
    """rpn.py:82-105: cross-entropy over the sampled anchors + smooth-L1 (sigma 3) on the box deltas.  The reference
    gathers the anchors with label != -1 first; ignore_index = -1 is the same mean over the same anchors."""
    if rpn.RPN_anchor_target is None:
        rpn.RPN_anchor_target = _AnchorTargetLayer(rpn.feat_stride, rpn.anchor_scales, rpn.anchor_ratios)
    B = rpn_cls_score.size(0)
    labels, targets, iw, ow = rpn.RPN_anchor_target((rpn_cls_score.detach(), gt_boxes[:, :, :5], im_info, num_boxes))
    score = rpn.reshape(rpn_cls_score, 2).permute(0, 2, 3, 1).contiguous().view(B, -1, 2)
    loss_cls = F.cross_entropy(score.view(-1, 2), labels.view(-1).long(), ignore_index=-1)
    loss_box = _smooth_l1_loss(rpn_bbox_pred, targets, iw, ow, sigma=3, dim=[1, 2, 3])
    return loss_cls, loss_box
