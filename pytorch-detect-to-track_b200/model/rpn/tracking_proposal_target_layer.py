"""Training-only tracking targets (lib/model/rpn/tracking_proposal_target_layer.py:20-196): the RoIs are the
ground-truth boxes of frame t; a box whose track id also occurs in frame t+tau gets the (normalised) box
regression from its frame-t box to its frame-t+tau box as target."""
import torch
import torch.nn as nn

from model.utils.config import cfg
from model.utils.net_utils import device_const
from .bbox_transform import bbox_transform_batch


class _TrackingProposalTargetLayer(nn.Module):
    def __init__(self, nclasses):
        super(_TrackingProposalTargetLayer, self).__init__()
        self._num_classes = nclasses

    @torch.no_grad()
    def forward(self, gt_boxes, num_boxes):
        """gt_boxes [2, B, K, 6] = (x1, y1, x2, y2, cls, track_id), num_boxes [2, B, 1].
        Batched over the images with static shapes and no device->host round trip (the reference loops over the images
        with ``nonzero`` and python branches, tracking_proposal_target_layer.py:60-120): the boxes whose track id occurs
        in both frames are moved to the front in track-id order by a stable sort, the rest are masked."""
        means = device_const(cfg.TRAIN.BBOX_NORMALIZE_MEANS, gt_boxes)
        stds = device_const(cfg.TRAIN.BBOX_NORMALIZE_STDS, gt_boxes)
        inside = device_const(cfg.TRAIN.BBOX_INSIDE_WEIGHTS, gt_boxes)
        B, K = gt_boxes.size(1), gt_boxes.size(2)
        dev = gt_boxes.device
        ar = torch.arange(K, device=dev).view(1, K)
        n0, n1 = num_boxes[0].view(B, 1).long(), num_boxes[1].view(B, 1).long()
        v0, v1 = ar < n0, ar < n1                                            # the real boxes of each frame [B, K]
        g0, g1 = gt_boxes[0], gt_boxes[1]
        corr = (g0[:, :, 5].unsqueeze(2) == g1[:, :, 5].unsqueeze(1)) & v0.unsqueeze(2) & v1.unsqueeze(1)   # N_t x N_t+tau
        rows, cols = corr.any(2), corr.any(1)
        cr, cc = rows.sum(1, keepdim=True), cols.sum(1, keepdim=True)
        valid = (n0 > 0) & (n1 > 0) & (cr > 0) & (cc > 0)                    # [B, 1]  (the reference `continue`s otherwise)
        big = torch.finfo(gt_boxes.dtype).max

        def aligned(g, sel, cnt):
            order = torch.sort(torch.where(sel, g[:, :, 5], torch.full_like(g[:, :, 5], big)), dim=1, stable=True)[1]
            out = g.gather(1, order.unsqueeze(2).expand(B, K, 6))
            return out * ((ar < cnt) & valid).unsqueeze(2).to(g.dtype)       # align the tracks across the frames

        t0, t1 = aligned(g0, rows, cr), aligned(g1, cols, cc)
        labels = t0[:, :, 4].clone()
        rois = gt_boxes.new_zeros(B, K, 5)
        rois[:, :, 0] = torch.arange(B, device=dev, dtype=rois.dtype).view(B, 1)
        rois[:, :, 1:] = g0[:, :, :4]
        rois = rois * valid.unsqueeze(2).to(rois.dtype)
        targets = (bbox_transform_batch(t0[:, :, :4], t1[:, :, :4]) - means) / stds
        fg = (labels > 0).unsqueeze(2).float()
        inside_w = inside.view(1, 1, 4) * fg
        return rois, labels, targets * fg, inside_w, (inside_w > 0).float()
