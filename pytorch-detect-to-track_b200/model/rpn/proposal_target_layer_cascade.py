"""Training-only RoI sampling (lib/model/rpn/proposal_target_layer_cascade.py:20-208) and the training branch
of the D&T graph (lib/model/faster_rcnn/rfcn.py:113-160, 176-204), restated for Python 3 / current torch.
Host-side training glue; sampling uses torch's generator on the device; no device->host round trip."""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from model.utils.config import cfg
from model.utils.net_utils import _smooth_l1_loss, device_const
from .bbox_transform import bbox_overlaps_batch, bbox_transform_batch


class _ProposalTargetLayer(nn.Module):
    """Assign proposals to ground truth: labels + (normalised) regression targets for 128 sampled RoIs / image.
    Static shapes, no device->host round trip (the head of the training step replays as a CUDA graph): the reference's
    per-image ``nonzero`` / ``randperm`` / python branches (proposal_target_layer_cascade.py:129-187) become batched sorts
    -- foreground: a random order of the candidates (random keys, non-candidates last), the first min(32, n_fg) taken;
    background (and foreground when there is no background): drawn WITH replacement from the compacted candidate list,
    floor(u * n) like the reference."""

    def __init__(self, nclasses):
        super(_ProposalTargetLayer, self).__init__()
        self._num_classes = nclasses
        self.generator = None

    @torch.no_grad()
    def forward(self, all_rois, gt_boxes, num_boxes):
        means = device_const(cfg.TRAIN.BBOX_NORMALIZE_MEANS, gt_boxes)
        stds = device_const(cfg.TRAIN.BBOX_NORMALIZE_STDS, gt_boxes)
        inside = device_const(cfg.TRAIN.BBOX_INSIDE_WEIGHTS, gt_boxes)
        gt_append = gt_boxes.new_zeros(gt_boxes.size(0), gt_boxes.size(1), 5)
        gt_append[:, :, 1:5] = gt_boxes[:, :, :4]
        all_rois = torch.cat([all_rois, gt_append], 1)                      # ground truth joins the candidates
        R = int(cfg.TRAIN.BATCH_SIZE)
        fg_per_image = max(1, int(np.round(cfg.TRAIN.FG_FRACTION * R)))
        overlaps = bbox_overlaps_batch(all_rois, gt_boxes[:, :, :5])
        max_ov, assign = overlaps.max(2)
        B, N = max_ov.shape
        dev = max_ov.device
        labels = torch.gather(gt_boxes[:, :, 4], 1, assign)
        fg_mask = max_ov >= cfg.TRAIN.FG_THRESH
        bg_mask = (max_ov < cfg.TRAIN.BG_THRESH_HI) & (max_ov >= cfg.TRAIN.BG_THRESH_LO)
        nf, nb = fg_mask.sum(1), bg_mask.sum(1)                             # [B]
        # (the reference raises when an image has neither; that cannot be tested without a host round trip: such a row
        # falls through to index 0 with background labels)
        keys = torch.rand(B, N, device=dev, generator=self.generator)
        fg_order = torch.sort(torch.where(fg_mask, keys, torch.full_like(keys, 2.0)), dim=1)[1]      # random order, candidates first
        fg_list = torch.sort((~fg_mask).to(torch.uint8), dim=1, stable=True)[1]                      # index order, candidates first
        bg_list = torch.sort((~bg_mask).to(torch.uint8), dim=1, stable=True)[1]
        u = torch.rand(B, R, device=dev, generator=self.generator)
        pick = lambda lst, n: lst.gather(1, torch.floor(u * n.view(B, 1).float()).long().clamp(min=0, max=N - 1))
        has_f, has_b = (nf > 0).view(B, 1), (nb > 0).view(B, 1)
        n_fg = torch.where(has_b.view(B), nf.clamp(max=fg_per_image), torch.full_like(nf, R)) * has_f.view(B).long()
        fg_rows = torch.where(has_b, fg_order[:, :R], pick(fg_list, nf))    # without replacement / (no background) with
        keep = torch.where(torch.arange(R, device=dev).view(1, R) < n_fg.view(B, 1), fg_rows, pick(bg_list, nb))
        is_fg_row = torch.arange(R, device=dev).view(1, R) < n_fg.view(B, 1)
        labels_b = labels.gather(1, keep) * is_fg_row.to(labels.dtype)      # background label 0
        rois_b = all_rois.gather(1, keep.unsqueeze(2).expand(B, R, 5)).clone()
        rois_b[:, :, 0] = torch.arange(B, device=dev, dtype=rois_b.dtype).view(B, 1)
        gt_b = gt_boxes.gather(1, assign.gather(1, keep).unsqueeze(2).expand(B, R, gt_boxes.size(2)))
        targets = (bbox_transform_batch(rois_b[:, :, 1:5], gt_b[:, :, :4]) - means) / stds
        fg_m = (labels_b > 0).unsqueeze(2).float()
        bbox_targets = targets * fg_m
        inside_w = inside.view(1, 1, 4) * fg_m
        outside_w = (inside_w > 0).float()
        return rois_b, labels_b, bbox_targets, inside_w, outside_w


def train_heads(net, B, conv3, conv4, conv5, base_feat, rfcn_cls, rfcn_bbox, info, gt_boxes, num_boxes, rpn_maps=None,
                trk_map=None):
    """Training branch of _RFCN.forward.  Tensors are leg-major over 2B images; gt_boxes [B,2,K,6],
    num_boxes [B,2,1].  Returns the reference's 10-tuple (rfcn.py:248-250).
    rpn_maps = (rpn_cls_score, rpn_bbox_pred) [2B, ...] and trk_map [B, 4*n_reg*49, H, W]: the outputs of the RPN /
    tracking-head convolutions when the caller has already run them (d2t_b200.train: the tcgen05 engine); base_feat and
    conv3/4/5 are then unused."""
    from .tracking_proposal_target_layer import _TrackingProposalTargetLayer
    L = 2
    if net.RFCN_proposal_target is None:
        net.RFCN_proposal_target = _ProposalTargetLayer(net.n_classes)
        net.RFCN_tracking_proposal_target = _TrackingProposalTargetLayer(net.n_classes)
    gt = gt_boxes.permute(1, 0, 2, 3).contiguous()          # [2, B, K, 6]
    nb = num_boxes.permute(1, 0, 2).contiguous()
    rois, rois_label, cls_prob, bbox_pred = [], [], [], []
    l_rpn_cls, l_rpn_box, l_cls, l_box = [], [], [], []
    all_rois = None
    if rpn_maps is not None:
        # the proposal step of both legs in one pass over the 2B images (rfcn.py:104-105 runs it per leg): same proposals,
        # half the launches, and the sequential NMS sweeps of all images side by side
        with torch.no_grad():
            all_rois = net.RFCN_rpn.proposals_from_maps(rpn_maps[0], rpn_maps[1], info)
            all_rois = torch.cat([all_rois[..., :1] - (torch.arange(L * B, device=all_rois.device) // B * B).view(-1, 1, 1).to(all_rois.dtype),
                                  all_rois[..., 1:]], dim=-1)                   # image index inside the leg
    for leg in range(L):
        sl = slice(leg * B, (leg + 1) * B)
        if rpn_maps is not None:
            leg_rois, lc, lb = net.RFCN_rpn.forward_from_maps(rpn_maps[0][sl], rpn_maps[1][sl], info[sl], gt[leg][:, :, :5],
                                                              nb[leg], rois=all_rois[sl])
        else:
            leg_rois, lc, lb = net.RFCN_rpn(base_feat[sl], info[sl], gt[leg][:, :, :5], nb[leg])
        l_rpn_cls.append(lc.view(1)), l_rpn_box.append(lb.view(1))
        leg_rois, label, target, iw, ow = net.RFCN_proposal_target(leg_rois, gt[leg][:, :, :5], nb[leg])
        label = label.view(-1).long()
        flat = leg_rois.view(-1, 5)
        pooled_cls = net.RFCN_psroi_cls_pool(rfcn_cls[sl].contiguous(), flat)
        pooled_loc = net.RFCN_psroi_loc_pool(rfcn_bbox[sl].contiguous(), flat)
        score = net.RFCN_cls_score(pooled_cls).view(flat.size(0), -1)
        pred = net.RFCN_bbox_pred(pooled_loc).view(flat.size(0), -1)
        if not net.class_agnostic:
            pred = torch.gather(pred.view(pred.size(0), -1, 4), 1, label.view(-1, 1, 1).expand(-1, 1, 4)).squeeze(1)
        l_cls.append(F.cross_entropy(score, label).view(1))
        l_box.append(_smooth_l1_loss(pred, target.view(-1, target.size(2)), iw.view(-1, iw.size(2)),
                                     ow.view(-1, ow.size(2))).view(1))
        rois.append(leg_rois), rois_label.append(label)
        cls_prob.append(F.softmax(score, dim=1).view(B, leg_rois.size(1), -1))
        bbox_pred.append(pred.view(B, leg_rois.size(1), -1))
    # ---- tracking branch
    trk = trk_map if trk_map is not None else net._tracking_maps(conv3, conv4, conv5, rfcn_bbox, B)
    t_rois, t_label, t_target, t_iw, t_ow = net.RFCN_tracking_proposal_target(gt, nb)
    pooled = net.RFCN_psroi_loc_pool(trk, t_rois.contiguous().view(-1, 5))
    tracking_pred = net.RFCN_tracking_pred(pooled).view(-1, pooled.size(1))
    l_trk = _smooth_l1_loss(tracking_pred, t_target.view(-1, t_target.size(2)), t_iw.view(-1, t_iw.size(2)),
                            t_ow.view(-1, t_ow.size(2)))
    return (torch.stack(rois), torch.stack(cls_prob), torch.stack(bbox_pred), tracking_pred, torch.stack(l_rpn_cls),
            torch.stack(l_rpn_box), torch.stack(l_cls), torch.stack(l_box), torch.stack(rois_label).view(L, B, -1), l_trk)
