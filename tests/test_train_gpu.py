"""One training step of the D&T graph on the GPU through the reference-shaped modules: target layers, the
five losses, backward through the PSRoI / correlation kernels' own backward kernels, SGD update."""
import numpy as np
import pytest
import torch

import common

pytestmark = pytest.mark.gpu


def test_training_step_runs_and_learns():
    from model.faster_rcnn.resnet import resnet
    torch.manual_seed(3)
    net = resnet(tuple(range(31)), 50, class_agnostic=True).create_architecture().cuda()
    net.train()
    B, H, W = 2, 224, 320
    g = torch.Generator().manual_seed(1)
    im_data = (torch.rand(B, 2, 3, H, W, generator=g) * 2 - 1).cuda()
    im_info = torch.tensor([H, W, 1.0]).view(1, 1, 3).expand(B, 2, 3).contiguous().cuda()
    gt = torch.from_numpy(common.make_gt_boxes(B, 30, seed=2, height=H, width=W)).cuda()
    nb = (gt[..., 4] > 0).sum(-1, keepdim=True)
    params = [p for p in net.parameters() if p.requires_grad]
    opt = torch.optim.SGD(params, lr=1e-4, momentum=0.9)
    losses = []
    for it in range(3):
        out = net(im_data, im_info, gt, nb)
        rois, cls_prob, bbox_pred, tracking_pred, l_rpn_cls, l_rpn_box, l_cls, l_box, rois_label, l_trk = out
        assert rois.shape == (2, B, 128, 5) and cls_prob.shape == (2, B, 128, 31) and rois_label.shape == (2, B, 128)
        loss = l_rpn_cls.mean() + l_rpn_box.mean() + l_cls.mean() + l_box.mean() + l_trk.mean()   # trainval_net.py:367-368
        assert torch.isfinite(loss)
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(float(loss))
    # gradients reached the trunk through the correlation and PSRoI backward kernels
    assert net.RFCN_base[6][0].conv1.weight.grad is not None and float(net.RFCN_base[6][0].conv1.weight.grad.abs().sum()) > 0
    assert float(net.corr_bbox_net.weight.grad.abs().sum()) > 0
    assert net.RFCN_base[4][0].conv1.weight.grad is None          # frozen stem / layer1 (resnet.py:279-289)
    assert losses[-1] < losses[0], losses
