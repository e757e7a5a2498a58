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
