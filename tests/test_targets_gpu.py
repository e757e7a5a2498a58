"""The three training target layers ON THE DEVICE against the outputs of the reference's own Python
(tests/golden/targets_reference.npz, made by tests/golden/make_golden_targets.py), and on the config-2 ground-truth
generator against the same layers run on the CPU.  Random subsampling differs by construction between the reference
(numpy RNG on the host) and these layers (torch generator on the device): deterministic quantities are compared exactly,
sampled ones through their invariants."""
import os

import numpy as np
import pytest
import torch

import common
from model.rpn.anchor_target_layer import _AnchorTargetLayer
from model.rpn.proposal_target_layer_cascade import _ProposalTargetLayer
from model.rpn.tracking_proposal_target_layer import _TrackingProposalTargetLayer
from model.utils.config import cfg

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(common.GOLDEN, "targets_reference.npz"))


def test_anchor_and_tracking_targets_on_device_match_reference(gold):
    gt = gold["gt"]
    layer = _AnchorTargetLayer(16, cfg.ANCHOR_SCALES, cfg.ANCHOR_RATIOS)
    g = torch.from_numpy(gt[:, 0, :, :5].copy()).cuda()
    nb = torch.from_numpy((gt[:, 0, :, 4] > 0).sum(1).astype(np.int64)).view(2, 1).cuda()
    info = torch.tensor([[300., 500., 1.], [300., 500., 1.]]).cuda()
    lab, tgt, iw, ow = [t.cpu().numpy() for t in layer((torch.zeros(2, 24, 19, 32, device="cuda"), g, info, nb))]
    ref_lab = gold["at_labels"]
    np.testing.assert_array_equal(lab == 1, ref_lab == 1)                       # positives: below the quota, identical sets
    for b in range(2):
        assert (lab[b] >= 0).sum() == 256 == (ref_lab[b] >= 0).sum()
    np.testing.assert_allclose(tgt, gold["at_targets"], rtol=1e-5, atol=1e-6)
    np.testing.assert_array_equal(iw, gold["at_iw"])
    gt_l = torch.from_numpy(gt.transpose(1, 0, 2, 3).copy()).cuda()
    nb_l = torch.from_numpy((gt.transpose(1, 0, 2, 3)[..., 4] > 0).sum(-1).astype(np.int64)).view(2, 2, 1).cuda()
    r, l, t, tiw, tow = [x.cpu().numpy() for x in _TrackingProposalTargetLayer(31)(gt_l, nb_l)]
    np.testing.assert_array_equal(r, gold["trk_rois"])
    np.testing.assert_array_equal(l, gold["trk_labels"])
    np.testing.assert_allclose(t, gold["trk_targets"], rtol=1e-5, atol=1e-6)
    np.testing.assert_array_equal(tiw, gold["trk_iw"])
    np.testing.assert_array_equal(tow, gold["trk_ow"])


def test_target_layers_device_equals_host_on_config2_ground_truth():
    """config-2 generator (600x1000, 1-5 tracked boxes per frame): the same layers on cuda and on the host -- identical
    deterministic outputs; the RoI sampler's invariants (128 RoIs, <= 32 foreground, labels / targets consistent)"""
    B = 2
    gt = common.make_gt_boxes(B, 30, seed=2, height=600, width=1000)             # [B, 2, 30, 6]
    gt_l = torch.from_numpy(gt.transpose(1, 0, 2, 3).copy())
    nb_l = torch.from_numpy((gt.transpose(1, 0, 2, 3)[..., 4] > 0).sum(-1).astype(np.int64)).view(2, B, 1)
    host = _TrackingProposalTargetLayer(31)(gt_l, nb_l)
    dev = _TrackingProposalTargetLayer(31)(gt_l.cuda(), nb_l.cuda())
    for a, b in zip(host, dev):
        np.testing.assert_allclose(a.numpy(), b.cpu().numpy(), rtol=1e-6, atol=1e-6)
    assert float((host[1] > 0).sum()) >= B                                        # tracks present in both frames
    # anchor targets: positives and regression targets are deterministic
    layer_h = _AnchorTargetLayer(16, cfg.ANCHOR_SCALES, cfg.ANCHOR_RATIOS)
    layer_d = _AnchorTargetLayer(16, cfg.ANCHOR_SCALES, cfg.ANCHOR_RATIOS)
    g0 = gt_l[0][:, :, :5].contiguous()
    info = torch.tensor([[600., 1000., 1.]] * B)
    score = torch.zeros(B, 24, 38, 63)
    lh = layer_h((score, g0, info, nb_l[0]))
    ld = layer_d((score.cuda(), g0.cuda(), info.cuda(), nb_l[0].cuda()))
    n_pos = int((lh[0] == 1).sum())
    if n_pos < 128 * B:                                                          # not subsampled: identical positive sets
        np.testing.assert_array_equal((lh[0] == 1).numpy(), (ld[0] == 1).cpu().numpy())
    np.testing.assert_allclose(lh[1].numpy(), ld[1].cpu().numpy(), rtol=1e-5, atol=1e-6)
    for b in range(B):
        assert int((ld[0][b] >= 0).sum()) == 256
    # proposal targets on the device
    rois = torch.from_numpy(np.stack([common.make_rois(2000, 1, 600, 1000, seed=11 + i, lo=30, hi=400) for i in range(B)])).cuda()
    layer = _ProposalTargetLayer(31)
    r, lab, tgt, iw, ow = layer(rois, g0.cuda(), nb_l[0].cuda())
    assert r.shape == (B, 128, 5) and lab.shape == (B, 128) and tgt.shape == (B, 128, 4)
    for b in range(B):
        assert int((lab[b] > 0).sum()) <= 32 and bool((r[b, :, 0] == b).all())
        fg = lab[b] > 0
        assert bool((iw[b][fg] == 1).all()) and bool((iw[b][~fg] == 0).all())
        assert bool((tgt[b][~fg] == 0).all())
        assert bool(torch.isin(lab[b][fg], g0[b, :, 4].cuda()).all())            # foreground labels are ground-truth classes


def test_rpn_proposals_of_both_legs_in_one_pass_equal_per_leg():
    """train_heads runs the proposal step ONCE over the 2B images of both legs (model/rpn/rpn.py: proposals_from_maps) where
    the reference runs the RPN per leg (rfcn.py:104-105): the images are independent, so the per-leg proposals must come out
    bit for bit, with the image index counted inside the leg."""
    from model.rpn.rpn import _RPN
    torch.manual_seed(7)
    B, L, H, W = 2, 2, 38, 63
    rpn = _RPN(512).cuda().train()
    score = torch.randn(L * B, rpn.nc_score_out, H, W, device="cuda")
    delta = torch.randn(L * B, rpn.nc_bbox_out, H, W, device="cuda") * 0.3
    info = torch.tensor([600.0, 1000.0, 1.0], device="cuda").repeat(L * B, 1)
    with torch.no_grad():
        both = rpn.proposals_from_maps(score, delta, info)
        for leg in range(L):
            sl = slice(leg * B, (leg + 1) * B)
            one = rpn.proposals_from_maps(score[sl], delta[sl], info[sl])
            assert torch.equal(both[sl][..., 1:], one[..., 1:])
            assert torch.equal(both[sl][..., 0] - leg * B, one[..., 0])
