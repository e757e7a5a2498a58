"""Training-only tracking targets (lib/model/rpn/tracking_proposal_target_layer.py:20-196): the RoIs are the
ground-truth boxes of frame t; a box whose track id also occurs in frame t+tau gets the (normalised) box
regression from its frame-t box to its frame-t+tau box as target."""
import torch
import torch.nn as nn

from model.utils.config import cfg
from .bbox_transform import bbox_transform_batch


class _TrackingProposalTargetLayer(nn.Module):
    def __init__(self, nclasses):
        super(_TrackingProposalTargetLayer, self).__init__()
        self._num_classes = nclasses

    @torch.no_grad()
    def forward(self, gt_boxes, num_boxes):
        """gt_boxes [2, B, K, 6] = (x1, y1, x2, y2, cls, track_id), num_boxes [2, B, 1]."""
        means = gt_boxes.new_tensor(cfg.TRAIN.BBOX_NORMALIZE_MEANS)
        stds = gt_boxes.new_tensor(cfg.TRAIN.BBOX_NORMALIZE_STDS)
        inside = gt_boxes.new_tensor(cfg.TRAIN.BBOX_INSIDE_WEIGHTS)
        B, K = gt_boxes.size(1), gt_boxes.size(2)
        t0 = gt_boxes.new_zeros(B, K, 6)
        t1 = gt_boxes.new_zeros(B, K, 6)
        labels = gt_boxes.new_zeros(B, K)
        rois = gt_boxes.new_zeros(B, K, 5)
        for b in range(B):
            n0, n1 = int(num_boxes[0][b][0]), int(num_boxes[1][b][0])
            if n0 == 0 or n1 == 0:
                continue
            g0, g1 = gt_boxes[0][b][:n0], gt_boxes[1][b][:n1]
            corr = g0[:, 5].view(-1, 1) == g1[:, 5].view(1, -1)              # N_t x N_t+tau
            rows = torch.nonzero(corr.sum(1)).view(-1)
            cols = torch.nonzero(corr.sum(0)).view(-1)
            if rows.numel() == 0 or cols.numel() == 0:
                continue
            a, c = g0[rows], g1[cols]
            a = a[torch.sort(a[:, 5])[1]]                                   # align the tracks across the frames
            c = c[torch.sort(c[:, 5])[1]]
            assert a.size(0) == c.size(0), "[tracking_proposal_target_layer] gt rois dim are not equal."
            t0[b, : a.size(0)] = a
            t1[b, : c.size(0)] = c
            labels[b] = t0[b][:, 4]
            rois[b, :, 0] = b
            rois[b, :, 1:] = gt_boxes[0][b][:, :4]
        targets = (bbox_transform_batch(t0[:, :, :4], t1[:, :, :4]) - means) / stds
        fg = (labels > 0).unsqueeze(2).float()
        inside_w = inside.view(1, 1, 4) * fg
        return rois, labels, targets * fg, inside_w, (inside_w > 0).float()
