"""End-to-end check of the D&T graph on the GPU (eval mode, small frames)."""
import numpy as np
import pytest
import torch

import common  # noqa: F401

pytestmark = pytest.mark.gpu


def _net(layers=50):
    from model.faster_rcnn.resnet import resnet
    torch.manual_seed(3)
    net = resnet(tuple(range(31)), layers, class_agnostic=True).create_architecture()
    return net.cuda().eval()


def test_rfcn_eval_forward_shapes_and_consistency(oracle):
    net = _net(50)
    B, H, W = 2, 224, 320
    g = torch.Generator().manual_seed(1)
    im_data = (torch.rand(B, 2, 3, H, W, generator=g) * 256 - 128).cuda()
    im_info = torch.tensor([H, W, 1.0]).view(1, 1, 3).expand(B, 2, 3).contiguous().cuda()
    with torch.no_grad():
        out = net(im_data, im_info, None, None)
    rois, cls_prob, bbox_pred, tracking_pred = out[:4]
    assert rois.shape == (2, B, 300, 5) and cls_prob.shape == (2, B, 300, 31)
    assert bbox_pred.shape == (2, B, 300, 4) and tracking_pred.shape == (B * 300, 4)
    for t in (rois, cls_prob, bbox_pred, tracking_pred):
        assert torch.isfinite(t).all()
    assert torch.allclose(cls_prob.sum(-1), torch.ones_like(cls_prob.sum(-1)), atol=1e-5)
    # col 0 is the per-leg image index (rfcn.py:108-112)
    assert set(rois[0, :, :, 0].unique().tolist()) <= {0.0, 1.0} and set(rois[1, :, :, 0].unique().tolist()) <= {0.0, 1.0}
    assert float(rois[..., 1:].min()) >= 0 and float(rois[..., 3].max()) <= W - 1 and float(rois[..., 4].max()) <= H - 1

    # the heads, recomputed on the CPU with the oracle ops from the network's own feature maps
    with torch.no_grad():
        frames = im_data.permute(1, 0, 2, 3, 4).reshape(2 * B, 3, H, W)
        conv3, conv4, conv5, base = net._im_to_head(frames)
        rfcn_cls, rfcn_bbox = net.RFCN_cls_net(base), net.RFCN_bbox_net(base)
        flat = rois.clone()
        flat[1, :, :, 0] += B
        flat = flat.view(-1, 5)
    pooled, _ = oracle.psroi_forward(rfcn_cls.cpu().numpy(), flat.cpu().numpy(), 1 / 16., 7, 7, 7, 31)
    score = torch.from_numpy(pooled.mean((2, 3)))
    np.testing.assert_allclose(cls_prob.view(-1, 31).cpu().numpy(), torch.softmax(score, 1).numpy(), rtol=1e-4, atol=1e-6)
    c4 = oracle.correlation_forward(conv4[:B].cpu().numpy(), conv4[B:].cpu().numpy(), 8, 1, 8, 1, 1)
    got = net.conv4_corr_layer(conv4[:B].contiguous(), conv4[B:].contiguous()).cpu().numpy()
    assert np.abs(got - c4).max() / np.abs(c4).max() < 1e-4
