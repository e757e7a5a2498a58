"""Detection decode + per-class NMS for every frame and class in ONE pass.

The reference does this after the network, on the host side of test_net.py:232-301: un-normalise the box
deltas, ``bbox_transform_inv_legs`` + ``clip_boxes``, divide by the image scale, then a python loop over the
30 classes -- ``nonzero(score > thresh)``, sort, ``cat``, ``nms(cls_dets, cfg.TEST.NMS)`` with a host round
trip per class -- and finally the ``max_per_image`` cut over all classes.  Here the (frame, class) pairs are
the batch axis of the on-device NMS (``d2t_nms_batched`` with per-list valid counts): one stable sort, one
gather, two kernel launches, no host synchronisation until the caller asks for python lists.

Keep-sets are the reference's: the same score-sorted lists go through the same greedy rule (IoU with +1
widths, strict ``>``), so they are bit-identical to calling ``nms`` class by class (tests/test_detect_gpu.py).
"""
import numpy as np
import torch

from model.rpn.bbox_transform import bbox_transform_inv_legs, clip_boxes
from . import ops


class Detections(object):
    """Device-resident result: ``dets [F, C-1, R, 5]`` (x1, y1, x2, y2, score; every list sorted by score, the first
    ``n_valid`` rows above the threshold), ``keep [F, C-1, R]`` int32 indices into the sorted list, ``num_keep [F, C-1]``.
    Frame f = leg * B + pair; class column j-1 holds class j (0 is background, test_net.py:273)."""

    def __init__(self, dets, n_valid, keep, num_keep, n_legs, n_pairs):
        self.dets, self.n_valid, self.keep, self.num_keep = dets, n_valid, keep, num_keep
        self.n_legs, self.n_pairs = n_legs, n_pairs

    def to_lists(self, max_per_image=0):
        """all_boxes[f][j] = float32 [K, 5] numpy arrays like test_net.py:283 (one device -> host copy), with the
        max_per_image cut of test_net.py:287-294 applied per frame."""
        dets, keep, num = self.dets.cpu().numpy(), self.keep.cpu().numpy(), self.num_keep.cpu().numpy()
        F, Cm1 = num.shape
        out = []
        for f in range(F):
            per = [np.zeros((0, 5), np.float32)]                     # class 0: background, never reported
            for c in range(Cm1):
                k = int(num[f, c])
                per.append(dets[f, c][keep[f, c, :k]] if k else np.zeros((0, 5), np.float32))
            if max_per_image > 0:
                scores = np.hstack([p[:, -1] for p in per[1:]])
                if len(scores) > max_per_image:
                    image_thresh = np.sort(scores)[-max_per_image]
                    per = [p[p[:, -1] >= image_thresh] if i else p for i, p in enumerate(per)]
            out.append(per)
        return out


def decode_boxes(rois, bbox_pred, im_info, stds, means):
    """test_net.py:236-262: rois [L, B, R, 5], bbox_pred [L, B, R, 4k], im_info [B, L, 3] -> boxes [L, B, R, 4k] in
    original-image pixels (divided by the frame scale)."""
    L, B, R, _ = rois.shape
    boxes = rois[..., 1:5]
    k4 = bbox_pred.size(-1)
    stds_t = torch.tensor(stds, dtype=torch.float32, device=rois.device).repeat(k4 // 4)
    means_t = torch.tensor(means, dtype=torch.float32, device=rois.device).repeat(k4 // 4)
    deltas = bbox_pred * stds_t + means_t
    pred = bbox_transform_inv_legs(boxes, deltas)
    info_lb = im_info.permute(1, 0, 2)                               # [L, B, 3]
    for l in range(L):
        pred[l] = clip_boxes(pred[l], info_lb[l])
    return pred / info_lb[..., 2].reshape(L, B, 1, 1)


def per_class_detections(rois, cls_prob, bbox_pred, im_info, thresh=0.0, nms_thresh=0.3, class_agnostic=True,
                         stds=(0.1, 0.1, 0.2, 0.2), means=(0.0, 0.0, 0.0, 0.0)):
    """rois [L, B, R, 5], cls_prob [L, B, R, C], bbox_pred [L, B, R, 4 or 4C], im_info [B, L, 3] (CUDA fp32) ->
    ``Detections`` for all L*B frames and the C-1 foreground classes (test_net.py:239-285 for every frame)."""
    L, B, R, C = cls_prob.shape
    F = L * B
    pred = decode_boxes(rois, bbox_pred, im_info, stds, means).reshape(F, R, -1)
    scores = cls_prob.reshape(F, R, C)[:, :, 1:].permute(0, 2, 1).contiguous()          # [F, C-1, R]
    valid = scores > thresh
    n_valid = valid.sum(-1).to(torch.int32)                                             # [F, C-1]
    # boxes at or below the threshold sort to the end of their list and are cut off by n_valid
    key = torch.where(valid, scores, torch.full_like(scores, -float("inf")))
    s_sorted, order = torch.sort(key, dim=-1, descending=True, stable=True)
    if class_agnostic:
        b = pred[:, None, :, :4].expand(F, C - 1, R, 4)
    else:
        b = pred.reshape(F, R, C, 4)[:, :, 1:].permute(0, 2, 1, 3)
    b = torch.gather(b, 2, order.unsqueeze(-1).expand(F, C - 1, R, 4))
    dets = torch.cat([b, torch.gather(scores, 2, order).unsqueeze(-1)], -1).contiguous()  # [F, C-1, R, 5]
    keep, num = ops.nms_batched(dets.reshape(F * (C - 1), R, 5), float(nms_thresh), n_valid=n_valid.reshape(-1).contiguous())
    return Detections(dets, n_valid, keep.reshape(F, C - 1, -1), num.reshape(F, C - 1), L, B)


def detect_reference_loop(rois, cls_prob, bbox_pred, im_info, thresh, nms_thresh, max_per_image):
    """test_net.py:232-294, frame by frame and class by class (one host round trip per class, like the reference)."""
    L, B, R, C = cls_prob.shape
    stds = torch.tensor((0.1, 0.1, 0.2, 0.2), device=rois.device)
    deltas = (bbox_pred.view(-1, 4) * stds).view(L, B, R, 4)
    pred = bbox_transform_inv_legs(rois[..., 1:5], deltas)
    info_lb = im_info.permute(1, 0, 2)
    for l in range(L):
        pred[l] = clip_boxes(pred[l], info_lb[l])
    pred = pred / info_lb[..., 2].reshape(L, B, 1, 1)
    out = []
    for l in range(L):
        for b in range(B):
            per = [np.zeros((0, 5), np.float32)]
            for j in range(1, C):
                inds = torch.nonzero(cls_prob[l, b, :, j] > thresh).view(-1)
                if inds.numel() > 0:
                    cls_scores = cls_prob[l, b][inds][:, j]
                    _, order = torch.sort(cls_scores, dim=0, descending=True, stable=True)
                    cls_dets = torch.cat([pred[l, b][inds, :], cls_scores.contiguous().view(-1, 1)], 1)[order]
                    keep = ops.nms(cls_dets.contiguous(), nms_thresh)
                    per.append(cls_dets[keep.view(-1).long()].cpu().numpy())
                else:
                    per.append(np.zeros((0, 5), np.float32))
            if max_per_image > 0:
                scores = np.hstack([p[:, -1] for p in per[1:]])
                if len(scores) > max_per_image:
                    t = np.sort(scores)[-max_per_image]
                    per = [p[p[:, -1] >= t] if i else p for i, p in enumerate(per)]
            out.append(per)
    return out


