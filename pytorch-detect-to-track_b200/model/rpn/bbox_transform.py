"""Box arithmetic of lib/model/rpn/bbox_transform.py used around the hot path:
bbox_transform_inv (:108-134), bbox_transform_inv_legs (:77-106), clip_boxes (:156-173).
Inside the proposal step these run fused in d2t_proposal_decode; the torch versions here serve
the detection decode after the network (test_net.py:239-262) and the tests.
"""
import torch


def bbox_transform_inv(boxes, deltas, batch_size=None):
    """boxes [B, N, 4], deltas [B, N, 4*k] -> [B, N, 4*k]; +1 width convention."""
    widths = boxes[:, :, 2] - boxes[:, :, 0] + 1.0
    heights = boxes[:, :, 3] - boxes[:, :, 1] + 1.0
    ctr_x = boxes[:, :, 0] + 0.5 * widths
    ctr_y = boxes[:, :, 1] + 0.5 * heights
    dx, dy, dw, dh = deltas[:, :, 0::4], deltas[:, :, 1::4], deltas[:, :, 2::4], deltas[:, :, 3::4]
    pcx = dx * widths.unsqueeze(2) + ctr_x.unsqueeze(2)
    pcy = dy * heights.unsqueeze(2) + ctr_y.unsqueeze(2)
    pw = torch.exp(dw) * widths.unsqueeze(2)
    ph = torch.exp(dh) * heights.unsqueeze(2)
    out = deltas.clone()
    out[:, :, 0::4] = pcx - 0.5 * pw
    out[:, :, 1::4] = pcy - 0.5 * ph
    out[:, :, 2::4] = pcx + 0.5 * pw
    out[:, :, 3::4] = pcy + 0.5 * ph
    return out


def bbox_transform_inv_legs(boxes, deltas, batch_size=None):
    """boxes / deltas carry a leading leg axis [n_legs, B, N, .] (bbox_transform.py:77-106)."""
    n_legs = boxes.size(0)
    return torch.stack([bbox_transform_inv(boxes[l], deltas[l]) for l in range(n_legs)], 0)


def clip_boxes(boxes, im_shape, batch_size=None):
    """Clamp x to [0, im_w-1], y to [0, im_h-1] with im_shape[i] = (h, w, scale)."""
    B = boxes.size(0)
    xmax = (im_shape[:, 1] - 1).view(B, 1, 1)
    ymax = (im_shape[:, 0] - 1).view(B, 1, 1)
    zero = torch.zeros_like(xmax)
    boxes[:, :, 0::4] = torch.min(torch.max(boxes[:, :, 0::4], zero), xmax)
    boxes[:, :, 1::4] = torch.min(torch.max(boxes[:, :, 1::4], zero), ymax)
    boxes[:, :, 2::4] = torch.min(torch.max(boxes[:, :, 2::4], zero), xmax)
    boxes[:, :, 3::4] = torch.min(torch.max(boxes[:, :, 3::4], zero), ymax)
    return boxes


def bbox_transform_batch(ex_rois, gt_rois):
    """Regression targets (dx, dy, dw, dh) of gt w.r.t. ex boxes, +1 width convention
    (bbox_transform.py:38-75).  ex_rois [N,4] or [B,N,4]; gt_rois [B,N,4]."""
    if ex_rois.dim() == 2:
        ex_rois = ex_rois.unsqueeze(0).expand_as(gt_rois[..., :4])
    ew = ex_rois[..., 2] - ex_rois[..., 0] + 1.0
    eh = ex_rois[..., 3] - ex_rois[..., 1] + 1.0
    ecx = ex_rois[..., 0] + 0.5 * ew
    ecy = ex_rois[..., 1] + 0.5 * eh
    gw = gt_rois[..., 2] - gt_rois[..., 0] + 1.0
    gh = gt_rois[..., 3] - gt_rois[..., 1] + 1.0
    gcx = gt_rois[..., 0] + 0.5 * gw
    gcy = gt_rois[..., 1] + 0.5 * gh
    return torch.stack(((gcx - ecx) / ew, (gcy - ecy) / eh, torch.log(gw / ew), torch.log(gh / eh)), -1)


def bbox_overlaps_batch(anchors, gt_boxes):
    """IoU (+1 convention) of anchors [N,4] / [B,N,4|5] against gt_boxes [B,K,>=4] -> [B,N,K]; all-zero
    (padding) gt boxes give 0, degenerate anchors give -1 (bbox_transform.py:208-298)."""
    B = gt_boxes.size(0)
    if anchors.dim() == 2:
        anchors = anchors.unsqueeze(0).expand(B, anchors.size(0), 4)
    elif anchors.size(2) == 5:
        anchors = anchors[:, :, 1:5]
    gt = gt_boxes[:, :, :4]
    gx = gt[:, :, 2] - gt[:, :, 0] + 1
    gy = gt[:, :, 3] - gt[:, :, 1] + 1
    ax = anchors[:, :, 2] - anchors[:, :, 0] + 1
    ay = anchors[:, :, 3] - anchors[:, :, 1] + 1
    g_area = (gx * gy).unsqueeze(1)
    a_area = (ax * ay).unsqueeze(2)
    a, g = anchors.unsqueeze(2), gt.unsqueeze(1)
    iw = (torch.min(a[..., 2], g[..., 2]) - torch.max(a[..., 0], g[..., 0]) + 1).clamp(min=0)
    ih = (torch.min(a[..., 3], g[..., 3]) - torch.max(a[..., 1], g[..., 1]) + 1).clamp(min=0)
    inter = iw * ih
    ov = inter / (a_area + g_area - inter)
    ov = ov.masked_fill(((gx == 1) & (gy == 1)).unsqueeze(1), 0)
    ov = ov.masked_fill(((ax == 1) & (ay == 1)).unsqueeze(2), -1)
    return ov
