"""End-to-end checks of the D&T graph on the GPU (eval mode).

First test: small frames through the reference-shaped nn.Module against the CPU oracle.  The others: parity of the whole eval forward AT THE BENCHMARKED CONFIGURATION (BASELINE.json configs[1]: Res-101 D&T, 600x1000
frames, 2 frame-pairs per GPU) -- the numbers bench.py times are the numbers checked here.

The reference graph (lib/model/faster_rcnn/rfcn.py:66-250 over resnet.py:258-344) calls cuDNN fp32 through torch for
its convolutions, so the comparator is the same nn.Module run by torch in fp32 with TF32 off, and -- because cuDNN fp32
itself carries rounding error over ~105 layers -- float64 as the ground truth both are measured against.

Tolerance reading: "within 1e-4 rel of the reference" is checked two ways, both written here:
  * max-norm:     max|a - b| <= 1e-4 * max|b|                                  (relative to the tensor's scale)
  * elementwise:  |a - b| <= 1e-4 * |b| + 2e-5 * max|b|  for every element     (rtol on each value, atol for values
                                                                               that are small against the scale)
"""
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


H, W, PAIRS = 600, 1000, 2


def max_rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


def assert_close_rel(a, b, name, rtol=1e-4, atol_of_scale=2e-5):
    a, b = a.double(), b.double()
    scale = float(b.abs().max())
    assert max_rel(a, b) <= rtol, (name, "max-norm", max_rel(a, b))
    bad = (a - b).abs() > rtol * b.abs() + atol_of_scale * scale
    assert not bool(bad.any()), (name, "elementwise", int(bad.sum()), float((a - b).abs().max()), scale)


@pytest.fixture(scope="module")
def res101():
    from bench import build_net, make_inputs
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    net = build_net(101).cuda()                      # exactly bench.py's network (seed 3, random init, identity BN)
    im_data, im_info = make_inputs(PAIRS, seed=1)    # exactly bench.py's rank-0 inputs
    return net, im_data.cuda(), im_info.cuda()


def test_engine_res101_600x1000(res101):
    """D2TEngine(resnet101, 2 pairs, 600x1000), 3xFP16 mode, eager and as a CUDA-graph replay, against torch fp32 and
    float64 on the same weights and frames: conv3/4/5, base_feat and the R-FCN maps within 1e-4, >= 98 % identical
    proposals, heads within 1e-4 on the identical proposals."""
    from d2t_b200.engine import D2TEngine, GraphedEngine
    net, im_data, im_info = res101
    N = 2 * PAIRS
    eng = D2TEngine(net, PAIRS, H, W, passes=16, keep_features=True)
    out = eng(im_data, im_info)
    torch.cuda.synchronize()
    for t in out[:4]:
        assert bool(torch.isfinite(t).all())
    frames = im_data.permute(1, 0, 2, 3, 4).reshape(N, 3, H, W).contiguous()
    with torch.no_grad():
        conv3, conv4, conv5, base = net._im_to_head(frames)
        cls_map, bbox_map = net.RFCN_cls_net(base), net.RFCN_bbox_net(base)
        ref = net(im_data, im_info)
    got = {"conv3": eng.feat_nchw[5], "conv4": eng.feat_nchw[6], "conv5": eng.feat_nchw[7],
           "base_feat": eng.base_feat.to_nchw(), "cls_map": eng.cls_map, "bbox_map": eng.bbox_map}
    want = {"conv3": conv3, "conv4": conv4, "conv5": conv5, "base_feat": base, "cls_map": cls_map, "bbox_map": bbox_map}
    errs = {k: max_rel(got[k], want[k]) for k in got}
    print("Res-101 600x1000 B=2, engine (3xFP16) vs torch fp32 graph, max rel err:", errs)
    for k in got:
        assert_close_rel(got[k], want[k], k)

    # ground truth: the trunk in float64, two frames at a time (fp64 cuDNN is slow; 4 frames of 600x1000 fit easily)
    net64 = __import__("copy").deepcopy(net).double()
    with torch.no_grad():
        f64 = [net64._im_to_head(frames[i:i + 1].double()) for i in range(N)]
    del net64
    c5_64 = torch.cat([f[2] for f in f64])
    base64 = torch.cat([f[3] for f in f64])
    e_eng = {"conv5": max_rel(got["conv5"], c5_64), "base_feat": max_rel(got["base_feat"], base64)}
    e_ref = {"conv5": max_rel(conv5, c5_64), "base_feat": max_rel(base, base64)}
    print("vs float64: engine", e_eng, " torch/cuDNN fp32", e_ref)
    assert max(e_eng.values()) < 1e-4, e_eng
    del f64, c5_64, base64

    # proposals: same boxes (coordinates to 1e-2 px; near-tied scores may swap places)
    same = (out[0] - ref[0]).abs().amax(-1) < 1e-2
    frac = float(same.float().mean())
    print("identical proposals: %.4f" % frac)
    assert frac >= 0.98, frac
    sel = same.view(-1)
    assert float((out[1].view(-1, 31)[sel] - ref[1].view(-1, 31)[sel]).abs().max()) < 1e-4       # softmax probabilities
    d = (out[2].view(-1, 4)[sel] - ref[2].view(-1, 4)[sel]).abs().max() / ref[2].abs().max()
    assert float(d) < 1e-4, float(d)
    sel0 = same[0].reshape(-1)
    d = (out[3][sel0] - ref[3][sel0]).abs().max() / ref[3].abs().max()
    assert float(d) < 1e-4, float(d)

    # the launch form bench.py times: CUDA-graph replay == the eager launches, bit for bit, replay after replay
    eager = [t.clone() for t in out[:4]]
    graphed = GraphedEngine(eng, PAIRS, H, W)
    for _ in range(3):
        rep = graphed(im_data, im_info)
        torch.cuda.synchronize()
        for a, b in zip(eager, rep[:4]):
            assert torch.equal(a, b)


def test_engine_res101_calibrated_bn(res101):
    """The same configuration with trained-looking BatchNorm statistics (activations O(1) instead of 1e7).  This random
    net is ILL-CONDITIONED: re-normalising every layer on two frames makes it amplify rounding noise from layer to layer
    (torch's own fp32 and fp64 forwards differ by ~1e-4 at conv5), so "1e-4 of the reference" is not a meaningful bar
    between two fp32 implementations here.  float64 is the arbiter: the engine's distance to float64 must stay within a
    small multiple of torch/cuDNN fp32's own distance (3xFP16 carries 22 mantissa bits against fp32's 24)."""
    import copy
    from d2t_b200.engine import D2TEngine
    from d2t_b200.synth import calibrate_batchnorm
    net0, im_data, im_info = res101
    net = copy.deepcopy(net0)
    N = 2 * PAIRS
    frames = im_data.permute(1, 0, 2, 3, 4).reshape(N, 3, H, W).contiguous()
    calibrate_batchnorm(net, frames)
    eng = D2TEngine(net, PAIRS, H, W, passes=16, keep_features=True)
    out = eng(im_data, im_info)
    for t in out[:4]:
        assert bool(torch.isfinite(t).all())
    with torch.no_grad():
        conv3, conv4, conv5, base = net._im_to_head(frames)
    net64 = copy.deepcopy(net).double()
    with torch.no_grad():
        f64 = [net64._im_to_head(frames[i:i + 1].double()) for i in range(N)]
    del net64
    truth = [torch.cat([f[j] for f in f64]) for j in range(4)]
    del f64
    got = (eng.feat_nchw[5], eng.feat_nchw[6], eng.feat_nchw[7], eng.base_feat.to_nchw())
    for name, a, b, t in zip(("conv3", "conv4", "conv5", "base_feat"), got, (conv3, conv4, conv5, base), truth):
        e_eng, e_ref, e_pair = max_rel(a, t), max_rel(b, t), max_rel(a, b)
        print("%s scale %.3g: engine vs fp64 %.2e, cuDNN fp32 vs fp64 %.2e, engine vs cuDNN %.2e" % (
            name, float(t.abs().max()), e_eng, e_ref, e_pair))
        assert e_eng < max(8 * e_ref, 2e-5), (name, e_eng, e_ref)
        assert e_eng < 1e-2, (name, e_eng)
