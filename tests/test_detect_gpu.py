"""Batched detection decode + per-class NMS (d2t_b200/detect.py) against the reference's own per-class loop
(test_net.py:239-301) restated literally on top of the single-list ``nms`` and against the CPU oracle."""
import numpy as np
import pytest
import torch

import common
from d2t_b200 import detect, ops
from model.rpn.bbox_transform import bbox_transform_inv_legs, clip_boxes

pytestmark = pytest.mark.gpu


def _inputs(L, B, R, C, seed, H=600., W=1000., scale=1.0, peaky=True):
    g = torch.Generator().manual_seed(seed)
    x1 = torch.rand(L, B, R, generator=g) * (W * 0.8)
    y1 = torch.rand(L, B, R, generator=g) * (H * 0.8)
    # clustered boxes so that NMS has something to suppress
    x1 = (x1 / 90).floor() * 90 + torch.rand(L, B, R, generator=g) * 25
    y1 = (y1 / 90).floor() * 90 + torch.rand(L, B, R, generator=g) * 25
    w = 40 + torch.rand(L, B, R, generator=g) * 200
    h = 40 + torch.rand(L, B, R, generator=g) * 200
    rois = torch.stack([torch.zeros(L, B, R), x1, y1, (x1 + w).clamp(max=W - 1), (y1 + h).clamp(max=H - 1)], -1)
    logits = torch.randn(L, B, R, C, generator=g) * (3.0 if peaky else 0.3)
    cls_prob = torch.softmax(logits, -1)
    bbox_pred = torch.randn(L, B, R, 4, generator=g)
    im_info = torch.tensor([H, W, scale]).view(1, 1, 3).expand(B, L, 3).contiguous()
    return rois.cuda(), cls_prob.cuda(), bbox_pred.cuda(), im_info.cuda()


_reference_loop = common.detect_reference_loop


@pytest.mark.parametrize("L,B,R,C,thresh,mpi", [(2, 2, 300, 31, 0.0, 100), (2, 1, 300, 31, 0.05, 0), (1, 3, 77, 5, 0.3, 10)])
def test_batched_detections_match_the_per_class_loop(L, B, R, C, thresh, mpi):
    rois, cls_prob, bbox_pred, im_info = _inputs(L, B, R, C, seed=40 + R)
    got = detect.per_class_detections(rois, cls_prob, bbox_pred, im_info, thresh=thresh, nms_thresh=0.3).to_lists(mpi)
    want = _reference_loop(rois, cls_prob, bbox_pred, im_info, thresh, 0.3, mpi)
    assert len(got) == len(want) == L * B
    for f in range(L * B):
        for j in range(C):
            np.testing.assert_array_equal(got[f][j], want[f][j], err_msg="frame %d class %d" % (f, j))


def test_batched_detections_keep_sets_vs_cpu_oracle(oracle):
    rois, cls_prob, bbox_pred, im_info = _inputs(1, 2, 300, 31, seed=7)
    d = detect.per_class_detections(rois, cls_prob, bbox_pred, im_info, thresh=0.01, nms_thresh=0.3)
    dets, nv, keep, num = d.dets.cpu().numpy(), d.n_valid.cpu().numpy(), d.keep.cpu().numpy(), d.num_keep.cpu().numpy()
    checked = 0
    for f in range(2):
        for c in range(30):
            n = int(nv[f, c])
            want = oracle.nms(dets[f, c, :n], 0.3) if n else np.zeros(0, np.int32)
            np.testing.assert_array_equal(keep[f, c, :int(num[f, c])], want)
            checked += n
    assert checked > 1000


def test_batched_detections_edge_cases():
    rois, cls_prob, bbox_pred, im_info = _inputs(1, 1, 50, 4, seed=3)
    # nothing above the threshold at all
    d = detect.per_class_detections(rois, cls_prob, bbox_pred, im_info, thresh=2.0)
    assert int(d.num_keep.sum()) == 0 and all(p.shape == (0, 5) for p in d.to_lists()[0])
    # one class empty, exact score ties in another, image scale != 1 (boxes are divided by it)
    cls_prob = cls_prob.clone()
    cls_prob[..., 1] = 0.0
    cls_prob[..., 2] = 0.25
    im_info2 = im_info.clone(); im_info2[..., 2] = 1.6
    got = detect.per_class_detections(rois, cls_prob, bbox_pred, im_info2, thresh=0.1).to_lists()
    want = _reference_loop(rois, cls_prob, bbox_pred, im_info2, 0.1, 0.3, 0)
    for j in range(4):
        np.testing.assert_array_equal(got[0][j], want[0][j])
    assert got[0][1].shape == (0, 5) and got[0][2].shape[0] > 0
    assert float(got[0][2][:, 2].max()) <= (1000 - 1) / 1.6 + 1e-3
