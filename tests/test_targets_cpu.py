"""Training-only target layers (host logic, CPU) against tests/golden/targets_reference.npz, produced by the
reference's own Python (tests/golden/make_golden_targets.py).  Random subsampling differs by construction
(numpy RNG on the host vs torch generator on the device), so sampled quantities are checked through their
invariants and everything deterministic exactly."""
import os

import numpy as np
import pytest
import torch

import common
from model.rpn.anchor_target_layer import _AnchorTargetLayer
from model.rpn.bbox_transform import bbox_overlaps_batch, bbox_transform_batch
from model.rpn.proposal_target_layer_cascade import _ProposalTargetLayer
from model.rpn.tracking_proposal_target_layer import _TrackingProposalTargetLayer
from model.utils.config import cfg


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(common.GOLDEN, "targets_reference.npz"))


def test_box_helpers_match_reference(gold):
    g = torch.from_numpy(gold["gt"][:, 0, :, :5].copy())
    rois = torch.from_numpy(common.make_rois(50, 1, 300, 500, seed=5)[:, 1:].copy())
    np.testing.assert_allclose(bbox_overlaps_batch(rois, g).numpy(), gold["ov_anchors2d"], rtol=1e-6, atol=1e-7)
    rois3 = torch.from_numpy(np.stack([common.make_rois(40, 1, 300, 500, seed=6 + i) for i in range(2)]))
    np.testing.assert_allclose(bbox_overlaps_batch(rois3, g).numpy(), gold["ov_rois3d"], rtol=1e-6, atol=1e-7)
    ex = rois3[:, :30, 1:5].contiguous()
    np.testing.assert_allclose(bbox_transform_batch(ex, g[:, :, :4] + 1.0).numpy(), gold["bt_batch"], rtol=1e-5, atol=1e-6)


def test_anchor_target_layer_vs_reference(gold):
    gt = gold["gt"]
    layer = _AnchorTargetLayer(16, cfg.ANCHOR_SCALES, cfg.ANCHOR_RATIOS)
    layer.generator = torch.Generator().manual_seed(0)
    g = torch.from_numpy(gt[:, 0, :, :5].copy())
    nb = torch.from_numpy((gt[:, 0, :, 4] > 0).sum(1).astype(np.int64)).view(2, 1)
    info = torch.tensor([[300., 500., 1.], [300., 500., 1.]])
    lab, tgt, iw, ow = [t.numpy() for t in layer((torch.zeros(2, 24, 19, 32), g, info, nb))]
    ref_lab = gold["at_labels"]
    assert lab.shape == ref_lab.shape
    # positives are below the fg quota here, so they are not subsampled: identical sets
    assert (ref_lab == 1).sum() < 128 * 2
    np.testing.assert_array_equal(lab == 1, ref_lab == 1)
    # 256 labelled anchors per image; negatives are a random subset of the same candidate set
    for b in range(2):
        assert (lab[b] >= 0).sum() == 256 == (ref_lab[b] >= 0).sum()
    np.testing.assert_allclose(tgt, gold["at_targets"], rtol=1e-5, atol=1e-6)          # deterministic
    np.testing.assert_array_equal(iw, gold["at_iw"])
    assert set(np.unique(ow).tolist()) == set(np.unique(gold["at_ow"]).tolist())        # 0 and 1/256


def test_tracking_target_layer_matches_reference(gold):
    gt = gold["gt"]
    gt_l = torch.from_numpy(gt.transpose(1, 0, 2, 3).copy())
    nb_l = torch.from_numpy((gt.transpose(1, 0, 2, 3)[..., 4] > 0).sum(-1).astype(np.int64)).view(2, 2, 1)
    r, l, t, iw, ow = _TrackingProposalTargetLayer(31)(gt_l, nb_l)
    np.testing.assert_array_equal(r.numpy(), gold["trk_rois"])
    np.testing.assert_array_equal(l.numpy(), gold["trk_labels"])
    np.testing.assert_allclose(t.numpy(), gold["trk_targets"], rtol=1e-5, atol=1e-6)
    np.testing.assert_array_equal(iw.numpy(), gold["trk_iw"])
    np.testing.assert_array_equal(ow.numpy(), gold["trk_ow"])
    assert (l.numpy() > 0).sum() >= 2          # the fixture does have cross-frame correspondences


def test_proposal_target_layer_invariants(gold):
    gt = torch.from_numpy(gold["gt"][:, 0, :, :5].copy())
    rois = torch.from_numpy(np.stack([common.make_rois(300, 1, 300, 500, seed=11 + i, lo=30, hi=300) for i in range(2)]))
    rois[1, :, 0] = 1
    layer = _ProposalTargetLayer(31)
    layer.generator = torch.Generator().manual_seed(1)
    out_rois, labels, targets, iw, ow = layer(rois, gt, None)
    assert out_rois.shape == (2, 128, 5) and labels.shape == (2, 128) and targets.shape == (2, 128, 4)
    ov = bbox_overlaps_batch(out_rois, gt)
    mx, arg = ov.max(2)
    for b in range(2):
        fg = labels[b] > 0
        assert 1 <= int(fg.sum()) <= 32                                    # 25 % of 128 at most (config.py:79-87)
        assert bool((mx[b][fg] >= 0.5).all()) and bool((mx[b][~fg] < 0.5).all())
        assert torch.equal(labels[b][fg], gt[b, arg[b][fg], 4])            # class of the best-overlapping gt box
        assert float(out_rois[b, :, 0].min()) == float(out_rois[b, :, 0].max()) == b
        want = bbox_transform_batch(out_rois[b:b + 1, :, 1:5], gt[b:b + 1, arg[b], :4])[0] / torch.tensor([0.1, 0.1, 0.2, 0.2])
        assert torch.allclose(targets[b][fg], want[fg], rtol=1e-5, atol=1e-6)
        assert float(targets[b][~fg].abs().max()) == 0.0
    assert torch.equal(ow, (iw > 0).float())


def test_anchor_target_layer_quota_with_many_positives():
    """More than 128 positives per image: the positives are cut to 128 and the negatives' quota is RPN_BATCHSIZE minus the
    number of positives BEFORE that cut (reference anchor_target_layer.py:119-137: `num_bg = RPN_BATCHSIZE - sum_fg[i]`),
    i.e. fewer than 128 negatives -- none at all once there were 256 positives."""
    H, W = 19, 32
    rng = np.random.RandomState(3)

    def boxes(n, size):
        x = rng.uniform(0, 500 - size, n)
        y = rng.uniform(0, 300 - size, n)
        return np.stack([x, y, x + size, y + size, np.ones(n)], 1).astype(np.float32)

    gt = np.zeros((2, 200, 5), np.float32)
    gt[0, :12] = boxes(12, 128)           # a handful of anchor-sized boxes: a few dozen positives
    gt[1, :200] = boxes(200, 128)         # crowded image: several hundred positives
    layer = _AnchorTargetLayer(16, cfg.ANCHOR_SCALES, cfg.ANCHOR_RATIOS)
    layer.generator = torch.Generator().manual_seed(0)
    g = torch.from_numpy(gt)
    info = torch.tensor([[300., 500., 1.], [300., 500., 1.]])
    lab = layer((torch.zeros(2, 24, H, W), g, info, None))[0].view(2, -1)
    # the positives / negatives before subsampling: the same rules with an unlimited quota
    saved = cfg.TRAIN.RPN_BATCHSIZE
    try:
        cfg.TRAIN.RPN_BATCHSIZE = 10 ** 6                      # no quota: every positive / negative survives
        full = layer((torch.zeros(2, 24, H, W), g, info, None))[0].view(2, -1)
    finally:
        cfg.TRAIN.RPN_BATCHSIZE = saved
    for b in range(2):
        n_pos_before = int((full[b] == 1).sum())
        n_pos, n_neg = int((lab[b] == 1).sum()), int((lab[b] == 0).sum())
        assert n_pos == min(128, n_pos_before)
        assert n_neg == max(0, 256 - n_pos_before), (n_pos_before, n_neg)
        assert bool(((lab[b] == 1) <= (full[b] == 1)).all()) and bool(((lab[b] == 0) <= (full[b] == 0)).all())   # subsets
    assert int((full[1] == 1).sum()) > 128                      # the fixture does exercise the cut
