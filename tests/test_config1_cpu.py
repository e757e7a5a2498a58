"""BASELINE.json configs[0] (SURVEY.md 8d "Config 1 -- plumbing"): Res-50 topology + RPN forward on one synthetic
300x500 frame, CPU-only PyTorch, no custom CUDA op.  Checks the graph wiring the GPU engine is built from -- the mirrored
``model.faster_rcnn.resnet`` module (resnet.py:66-129, 258-344) and the restated proposal step (proposal_layer.py:49-161,
pinned against the reference's own Python in tests/test_oracle_golden.py) -- at the shapes the survey lists:
base feature [1, 512, 19, 32], 19*32*12 = 7 296 anchors, 300 RoIs, 48.8 GFLOP."""
import numpy as np
import torch
import torch.nn.functional as F

import common  # noqa: F401  (sys.path)


def _conv_flops(net, x):
    total = [0.0]
    hooks = []

    def hook(m, inp, out):
        total[0] += 2.0 * out.numel() * m.in_channels * m.kernel_size[0] * m.kernel_size[1] / m.groups
    for m in net.modules():
        if isinstance(m, torch.nn.Conv2d):
            hooks.append(m.register_forward_hook(hook))
    return total, hooks


def test_config1_res50_rpn_forward_cpu(oracle):
    from model.faster_rcnn.resnet import resnet
    from model.utils.config import cfg
    torch.manual_seed(3)
    net = resnet(tuple(range(31)), 50, class_agnostic=True).create_architecture().eval()
    g = torch.Generator().manual_seed(0)
    frame = torch.randn(1, 3, 300, 500, generator=g) * 50.0
    im_info = np.array([[300.0, 500.0, 1.0]], np.float32)
    total, hooks = _conv_flops(net, frame)
    with torch.no_grad():
        conv3, conv4, conv5, base = net._im_to_head(frame)
        rpn = net.RFCN_rpn
        x = F.relu(rpn.RPN_Conv(base))
        score = rpn.RPN_cls_score(x)
        prob = rpn.cls_prob_from_score(score, rpn.nc_score_out)
        deltas = rpn.RPN_bbox_pred(x)
    for h in hooks:
        h.remove()
    # topology: stride on the first 1x1 of each stage, layer4 stride 1 / dilation 2, 3x3 dil-6 head
    assert tuple(conv3.shape) == (1, 512, 38, 63)
    assert tuple(conv4.shape) == (1, 1024, 19, 32)
    assert tuple(conv5.shape) == (1, 2048, 19, 32)
    assert tuple(base.shape) == (1, 512, 19, 32)
    A = len(cfg.ANCHOR_SCALES) * len(cfg.ANCHOR_RATIOS)
    assert A == 12 and tuple(prob.shape) == (1, 2 * A, 19, 32) and tuple(deltas.shape) == (1, 4 * A, 19, 32)
    assert 19 * 32 * A == 7296
    # pairwise softmax through the reshape (rpn.py:66-68): bg / fg of one anchor sum to 1
    p = prob.view(1, 2, A, 19, 32)
    assert float((p.sum(1) - 1).abs().max()) < 1e-6
    # the tracking head's input width (resnet.py:311): 2 x 196 loc maps + 81 + 289 + 289 correlation channels
    assert net.corr_bbox_net.in_channels == 1051
    # convolution work of the frame: SURVEY 8d quotes 48.8 GFLOP (trunk + 3x3 dil-6 head); the RPN's 3x3 512 -> 512 conv
    # and its two 1x1 heads add 2.9 GFLOP on the 19 x 32 map
    rpn_flops = 2.0 * 19 * 32 * 512 * (512 * 9 + 2 * A + 4 * A)
    assert abs((total[0] - rpn_flops) / 1e9 - 48.8) < 0.1, total[0] / 1e9
    anchors = oracle.generate_anchors(scales=tuple(cfg.ANCHOR_SCALES), ratios=tuple(cfg.ANCHOR_RATIOS)).astype(np.float32)
    rois = oracle.proposal_layer(prob.numpy(), deltas.numpy(), im_info, anchors, cfg.TEST.RPN_PRE_NMS_TOP_N,
                                 cfg.TEST.RPN_POST_NMS_TOP_N, cfg.TEST.RPN_NMS_THRESH)
    assert rois.shape == (1, 300, 5)
    assert float(np.abs(rois[..., 0]).max()) == 0.0                      # column 0 = image index
    b = rois[0, :, 1:]
    assert b[:, 0].min() >= 0 and b[:, 1].min() >= 0 and b[:, 2].max() <= 499 and b[:, 3].max() <= 299   # clip_boxes
    assert (b[:, 2] >= b[:, 0]).all() and (b[:, 3] >= b[:, 1]).all()
    # deterministic
    rois2 = oracle.proposal_layer(prob.numpy(), deltas.numpy(), im_info, anchors, cfg.TEST.RPN_PRE_NMS_TOP_N,
                                  cfg.TEST.RPN_POST_NMS_TOP_N, cfg.TEST.RPN_NMS_THRESH)
    np.testing.assert_array_equal(rois, rois2)
