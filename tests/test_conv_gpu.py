"""tcgen05 convolution engine against torch's fp32 conv2d (TF32 off) on the same inputs.
3-pass (3xTF32) mode must match to fp32 accuracy; 1-pass TF32 mode to ~1e-3."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import common  # noqa: F401
from d2t_b200 import conv as dc

pytestmark = pytest.mark.gpu


def _ref(x, w, scale, shift, stride, pad, dil, relu, res):
    torch.backends.cudnn.allow_tf32 = False
    y = F.conv2d(x.double(), w.double(), None, stride, pad, dil)
    if scale is not None:
        y = y * scale.double().view(1, -1, 1, 1)
    if shift is not None:
        y = y + shift.double().view(1, -1, 1, 1)
    if res is not None:
        y = y + res.double()
    return (F.relu(y) if relu else y).float()


CASES = [
    # N, Cin, H, W, Cout, k, stride, pad, dil, relu, residual
    (1, 64, 8, 16, 64, 1, 1, 0, 1, False, False),          # smallest: 1 tile, BN=64
    (2, 256, 38, 63, 64, 1, 1, 0, 1, True, False),         # layer1-style 1x1, W=63 (2x64 tiles)
    (2, 64, 38, 63, 256, 1, 1, 0, 1, True, True),          # expansion 1x1 + residual
    (2, 64, 38, 63, 64, 3, 1, 1, 1, True, False),          # 3x3 pad 1
    (1, 512, 38, 63, 512, 3, 1, 2, 2, True, False),        # layer4 dilated 3x3
    (1, 256, 19, 32, 512, 3, 1, 6, 6, True, False),        # dilation-6 head conv (bias)
    (2, 256, 75, 125, 128, 1, 2, 0, 1, True, False),       # stride-2 1x1 (first conv of a stage)
    (1, 64, 30, 250, 64, 1, 1, 0, 1, True, False),         # W > 128: two tiles per row
    (1, 512, 38, 63, 1519, 1, 1, 0, 1, False, False),      # R-FCN cls head: Cout not a multiple of 16
    (1, 1051, 19, 32, 196, 1, 1, 0, 1, False, False),      # tracking head: Cin padded to 1056
    (4, 256, 38, 63, 1024, 1, 1, 0, 1, True, True),        # layer3 closing 1x1 + residual: A-resident variant, 8 n tiles
    (2, 128, 75, 125, 512, 1, 1, 0, 1, True, True),        # layer2 closing 1x1 (2 K blocks)
    (2, 256, 75, 125, 512, 1, 2, 0, 1, False, False),      # stride-2 shortcut conv (A-resident, no residual)
    (1, 192, 19, 32, 300, 1, 1, 0, 1, True, False),        # 3 K blocks, Cout not a multiple of the tile
]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("passes", [16, 3, 1])
def test_conv_matches_torch(case, passes):
    N, Cin, H, W, Cout, k, stride, pad, dil, relu, use_res = case
    g = torch.Generator(device="cuda").manual_seed(1234 + Cin + Cout + k)
    x = torch.randn(N, Cin, H, W, device="cuda", generator=g)
    w = torch.randn(Cout, Cin, k, k, device="cuda", generator=g) * (2.0 / (Cin * k * k)) ** 0.5
    scale = torch.rand(Cout, device="cuda", generator=g) + 0.5
    shift = torch.randn(Cout, device="cuda", generator=g)
    OH = (H + 2 * pad - dil * (k - 1) - 1) // stride + 1
    OW = (W + 2 * pad - dil * (k - 1) - 1) // stride + 1
    res = torch.randn(N, Cout, OH, OW, device="cuda", generator=g) if use_res else None
    xs = dc.ActTensor.from_nchw(x)
    assert torch.equal(xs.to_nchw(Cin), x)
    rs = dc.ActTensor.from_nchw(res, cstride=Cout) if use_res else None
    layer = dc.ConvLayer(xs, w, scale, shift, stride, pad, dil, relu, rs, passes=passes, want_nhwc=(Cout % 4 == 0),
                         want_nchw=True)
    out = layer.run()
    torch.cuda.synchronize()
    want = _ref(x, w, scale, shift, stride, pad, dil, relu, res)
    got = layer.out_nchw
    assert got.shape == want.shape
    err = float((got - want).abs().max() / want.abs().max())
    print("case", case, "passes", passes, "max rel err %.2e" % err)
    tol = 1e-5 if passes != 1 else 3e-3
    assert err < tol, (err, layer.info)
    if layer.out is not None:
        assert torch.equal(layer.out.to_nchw(Cout), got)      # the NHWC output carries the same values


@pytest.mark.parametrize("xscale,wscale", [(3.0e7, 1.0), (1.0e-6, 1.0e-3), (1.0, 4.0e4)])
def test_fp16_split_conv_dynamic_range(xscale, wscale):
    """3xFP16: operands far outside fp16's range (the random-init Res-101 reaches 3e7) go through the per-tensor
    power-of-two scales; a chained second conv reads the first one's epilogue-written amax."""
    g = torch.Generator(device="cuda").manual_seed(99)
    x = torch.randn(2, 128, 38, 63, device="cuda", generator=g) * xscale
    x[0, 3, 5, 7] = 17.0 * xscale                                        # an outlier sets the scale
    w1 = torch.randn(256, 128, 3, 3, device="cuda", generator=g) * 0.03 * wscale
    w2 = torch.randn(64, 256, 1, 1, device="cuda", generator=g) * 0.06
    sh = torch.randn(256, device="cuda", generator=g) * xscale * wscale
    l1 = dc.ConvLayer(dc.ActTensor.from_nchw(x), w1, None, sh, 1, 1, 1, True, passes=16, want_nchw=True)
    l2 = dc.ConvLayer(l1.out, w2, None, None, passes=16, want_nchw=True)
    l1.run()
    l2.run()
    torch.cuda.synchronize()
    want1 = _ref(x, w1, None, sh, 1, 1, 1, True, None)
    want2 = _ref(want1, w2, None, None, 1, 0, 1, False, None)
    assert float(l1.out.amax) == float(l1.out_nchw.abs().max())          # the epilogue's running max is exact
    e1 = float((l1.out_nchw - want1).abs().max() / want1.abs().max())
    e2 = float((l2.out_nchw - want2).abs().max() / want2.abs().max())
    assert e1 < 1e-5 and e2 < 1e-5, (e1, e2)


def test_maxpool_ceil_mode():
    x = torch.randn(2, 64, 37, 50, device="cuda")
    xs = dc.ActTensor.from_nchw(x)
    out = dc.maxpool3x3s2(xs).to_nchw()
    want = F.max_pool2d(x, 3, 2, 0, ceil_mode=True)
    assert out.shape == want.shape and torch.equal(out, want)
    x = torch.randn(1, 64, 300, 500, device="cuda")
    out = dc.maxpool3x3s2(dc.ActTensor.from_nchw(x)).to_nchw()
    want = F.max_pool2d(x, 3, 2, 0, ceil_mode=True)
    assert out.shape == want.shape == (1, 64, 150, 250) and torch.equal(out, want)


@pytest.mark.parametrize("passes", [3, 16])
def test_stem_conv_7x7_stride2(passes):
    g = torch.Generator(device="cuda").manual_seed(5)
    for (N, H, W) in [(1, 64, 96), (2, 75, 101), (1, 600, 1000)]:
        x = torch.rand(N, 3, H, W, device="cuda", generator=g) * 256 - 128
        w = torch.randn(64, 3, 7, 7, device="cuda", generator=g) * (2.0 / (49 * 64)) ** 0.5
        scale = torch.rand(64, device="cuda", generator=g) + 0.5
        shift = torch.randn(64, device="cuda", generator=g)
        stem = dc.StemConv(N, H, W, w, scale, shift, relu=True, passes=passes)
        out = stem.run(x).to_nchw()
        want = _ref(x, w, scale, shift, 2, 3, 1, True, None)
        assert out.shape == want.shape
        err = float((out - want).abs().max() / want.abs().max())
        assert err < 1e-5, (err, N, H, W)


@pytest.mark.parametrize("passes", [16, 3])
def test_engine_matches_torch_graph(passes):
    """The whole eval forward on the sm_100a engine against the same nn.Module run by torch
    (cuDNN fp32, TF32 off) -- trunk features to ~1e-5, identical proposals, heads to 1e-4."""
    from model.faster_rcnn.resnet import resnet
    from d2t_b200.engine import D2TEngine
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(3)
    net = resnet(tuple(range(31)), 50, class_agnostic=True).create_architecture().cuda().eval()
    # non-trivial frozen BN statistics so the folding is exercised
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.1)
            m.running_var.uniform_(0.5, 1.5)
            m.weight.data.uniform_(0.5, 1.5)
            m.bias.data.normal_(0, 0.1)
    B, H, W = 2, 224, 320
    g = torch.Generator().manual_seed(1)
    im_data = (torch.rand(B, 2, 3, H, W, generator=g) * 256 - 128).cuda()
    im_info = torch.tensor([H, W, 1.0]).view(1, 1, 3).expand(B, 2, 3).contiguous().cuda()
    eng = D2TEngine(net, B, H, W, passes=passes, keep_features=True)
    out = eng(im_data, im_info)
    with torch.no_grad():
        frames = im_data.permute(1, 0, 2, 3, 4).reshape(2 * B, 3, H, W)
        conv3, conv4, conv5, base = net._im_to_head(frames)
        ref = net(im_data, im_info)

    def rel(a, b):
        return float((a - b).abs().max() / b.abs().max())

    errs = {"conv3": rel(eng.feat_nchw[5], conv3), "conv4": rel(eng.feat_nchw[6], conv4), "conv5": rel(eng.feat_nchw[7], conv5),
            "base": rel(eng.base_feat.to_nchw(), base), "cls_map": rel(eng.cls_map, net.RFCN_cls_net(base))}
    print("engine vs torch fp32 graph, max rel err:", errs)
    assert max(errs.values()) < 1e-4, errs          # north star: fp outputs within 1e-4 rel of the reference
    # proposals: same boxes (coordinates to 1e-2 px; the score order may swap near-ties)
    same = (out[0] - ref[0]).abs().amax(-1) < 1e-2
    assert float(same.float().mean()) > 0.98
    sel = same.view(-1)
    # heads, relative to each output's scale (random-init heads give large raw regression values)
    assert float((out[1].view(-1, 31)[sel] - ref[1].view(-1, 31)[sel]).abs().max()) < 1e-4
    d = (out[2].view(-1, 4)[sel] - ref[2].view(-1, 4)[sel]).abs().max() / ref[2].abs().max()
    assert float(d) < 1e-4, float(d)
    sel0 = same[0].reshape(-1)
    d = (out[3][sel0] - ref[3][sel0]).abs().max() / ref[3].abs().max()
    assert float(d) < 1e-4, float(d)


def test_engine_chains_and_graph_replay():
    """D2TEngineStreams (independent frame-pair chains on their own streams, enqueued in turn) computes what one
    D2TEngine per chain computes -- bit for bit, the kernels and their inputs are the same -- and agrees with the
    whole-batch engine to the fp32 bar; GraphedEngine replays it bit-identically, replay after replay (the stream-K
    hand-shake leaves no stale flag behind)."""
    from model.faster_rcnn.resnet import resnet
    from d2t_b200.engine import D2TEngine, D2TEngineStreams, GraphedEngine
    torch.manual_seed(3)
    net = resnet(tuple(range(31)), 50, class_agnostic=True).create_architecture().cuda().eval()
    B, H, W = 2, 224, 320
    g = torch.Generator().manual_seed(1)
    im_data = (torch.rand(B, 2, 3, H, W, generator=g) * 256 - 128).cuda()
    im_info = torch.tensor([H, W, 1.0]).view(1, 1, 3).expand(B, 2, 3).contiguous().cuda()
    whole = D2TEngine(net, B, H, W)(im_data, im_info)
    chains = D2TEngineStreams(net, B, H, W, chains=2)
    out = chains(im_data, im_info)
    torch.cuda.synchronize()
    single = D2TEngine(net, 1, H, W)
    for i in range(B):
        one = single(im_data[i:i + 1], im_info[i:i + 1])
        assert torch.equal(one[0][..., 1:], out[0][:, i:i + 1, :, 1:])          # rois (image index aside)
        assert float((out[0][:, i, :, 0] - i).abs().max()) == 0.0               # image index inside the whole batch
        assert torch.equal(one[1], out[1][:, i:i + 1])
        assert torch.equal(one[2], out[2][:, i:i + 1])
        R = one[3].size(0)
        assert torch.equal(one[3], out[3][i * R:(i + 1) * R])
    same = (out[0] - whole[0]).abs().amax(-1) < 1e-2
    assert float(same.float().mean()) > 0.98
    assert float((out[1][same] - whole[1][same]).abs().max()) < 1e-4
    graphed = GraphedEngine(chains, B, H, W)
    first = [t.clone() for t in graphed(im_data, im_info)[:4]]
    torch.cuda.synchronize()
    for a, b in zip(first, out[:4]):
        assert torch.equal(a, b)
    for _ in range(5):
        again = graphed(im_data, im_info)
        torch.cuda.synchronize()
        for a, b in zip(first, again[:4]):
            assert torch.equal(a, b)


@pytest.mark.parametrize("C,H,W,p,B", [(64, 20, 30, (8, 1, 8, 1, 1), 2), (1024, 38, 63, (8, 1, 8, 1, 1), 2),
                                       (512, 75, 125, (8, 1, 8, 2, 2), 1), (2048, 38, 63, (8, 1, 8, 1, 1), 1),
                                       (96, 21, 27, (4, 1, 4, 1, 1), 1), (40, 13, 50, (0, 1, 3, 1, 1), 1)])
@pytest.mark.parametrize("passes", [16, 3])
def test_tensor_core_correlation(C, H, W, p, B, passes, monkeypatch):
    """CORR mode of the tcgen05 kernel (3xFP16 and 3xTF32) against the reference kernel itself (or the fp32 SIMT kernel);
    inputs of different magnitudes: each operand carries its own scale in the fp16 split."""
    from d2t_b200 import ops
    from oracle import ref_cuda
    monkeypatch.setattr(ops, "CORRELATION_PASSES", passes)
    g = torch.Generator(device="cuda").manual_seed(77)
    a = torch.randn(B, C, H, W, device="cuda", generator=g) * 37.0
    b = torch.randn(B, C, H, W, device="cuda", generator=g) * 2.0e-3
    out = ops.correlation_forward(a, b, *p)
    ops.TENSOR_CORE_CORRELATION = False
    try:
        simt = ops.correlation_forward(a, b, *p)
    finally:
        ops.TENSOR_CORE_CORRELATION = True
    assert out.shape == simt.shape
    ref = ref_cuda.correlation_forward(a, b, *p) if ref_cuda.available() else simt
    err = float((out - ref).abs().max() / ref.abs().max())
    print("corr", (C, H, W, p), "max rel err vs reference kernel %.2e" % err)
    assert err < 2e-5, err
    assert float((simt - ref).abs().max() / ref.abs().max()) < 1e-4


def test_cta_pair_mode_matches(monkeypatch):
    """The cta_group::2 (CTA-pair) variant of the kernel gives the same result as single-CTA mode."""
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(2, 256, 38, 63, device="cuda", generator=g)
    w = torch.randn(384, 256, 3, 3, device="cuda", generator=g) * 0.02
    sc, sh = torch.rand(384, device="cuda", generator=g) + 0.5, torch.randn(384, device="cuda", generator=g)
    res = dc.ActTensor.from_nchw(torch.randn(2, 384, 38, 63, device="cuda", generator=g), cstride=384)
    outs = []
    for pair in ("0", "1"):
        monkeypatch.setenv("D2T_CONV_PAIR", pair)
        layer = dc.ConvLayer(dc.ActTensor.from_nchw(x), w, sc, sh, 1, 1, 1, True, res, passes=3, want_nchw=True)
        assert layer.info["grid"] % 10 == int(pair)
        layer.run()
        torch.cuda.synchronize()
        outs.append((layer.out_nchw.clone(), layer.out.x.clone()))
    want = _ref(x, w, sc, sh, 1, 1, 1, True, res.to_nchw())
    for o in outs:
        assert float((o[0] - want).abs().max() / want.abs().max()) < 1e-5
    assert float((outs[0][0] - outs[1][0]).abs().max() / want.abs().max()) < 2e-6
    assert torch.equal(outs[1][1].permute(0, 3, 1, 2), outs[1][0])


@pytest.mark.parametrize("case", [(2, 256, 38, 63, 384, 3, 1, 1, False), (4, 1024, 38, 63, 256, 1, 0, 1, False),
                                  (1, 512, 38, 63, 512, 3, 2, 2, False), (3, 256, 19, 32, 1519, 1, 0, 1, False),
                                  (2, 128, 38, 63, 512, 1, 0, 1, True)])
def test_cta_pair_mode_fp16_split(monkeypatch, case):
    """3xFP16 as CTA pairs (cta_group::2, M = 256: each CTA converts its own activation tile into its own tensor memory and
    stages HALF of the weight tile): the same products as single-CTA mode; only the stream-K split points differ (units are
    dealt to 74 pairs instead of 148 CTAs), i.e. the order of a few fp32 additions -- outputs agree to 2e-6 of the scale and
    both are within 1e-5 of float64.  Odd m-tile counts (the pair's second tile does not exist), stream-K splits across
    pairs, a residual, Cout not a multiple of the tile, NCHW + NHWC outputs, repeated launches."""
    N, Cin, H, W, Cout, k, pad, dil, use_res = case
    g = torch.Generator(device="cuda").manual_seed(21 + Cin + Cout)
    x = torch.randn(N, Cin, H, W, device="cuda", generator=g)
    w = torch.randn(Cout, Cin, k, k, device="cuda", generator=g) * (2.0 / (Cin * k * k)) ** 0.5
    sc, sh = torch.rand(Cout, device="cuda", generator=g) + 0.5, torch.randn(Cout, device="cuda", generator=g)
    res = dc.ActTensor.from_nchw(torch.randn(N, Cout, H, W, device="cuda", generator=g), cstride=Cout) if use_res else None
    outs = []
    for pair in ("0", "1"):
        monkeypatch.setenv("D2T_CONV_PAIR", pair)
        layer = dc.ConvLayer(dc.ActTensor.from_nchw(x), w, sc, sh, 1, pad, dil, True, res, passes=16, want_nhwc=(Cout % 4 == 0),
                             want_nchw=True)
        assert layer.info["grid"] % 10 == (int(pair) if Cout % 4 == 0 else 0)       # (pairs need the NHWC output path)
        for _ in range(2):
            layer.run()
        torch.cuda.synchronize()
        outs.append((layer.out_nchw.clone(), layer.out.x.clone() if layer.out is not None else None, layer.out.amax.clone() if layer.out is not None else None))
    want = _ref(x, w, sc, sh, 1, pad, dil, True, res.to_nchw() if use_res else None)
    scale = float(want.abs().max())
    assert float((outs[1][0] - want).abs().max()) / scale < 1e-5 and float((outs[0][0] - want).abs().max()) / scale < 1e-5
    assert float((outs[0][0] - outs[1][0]).abs().max()) / scale < 2e-6
    if outs[0][1] is not None:
        assert torch.equal(outs[1][1][..., :Cout].permute(0, 3, 1, 2), outs[1][0])      # NHWC and NCHW outputs carry the same values
        assert abs(float(outs[0][2]) - float(outs[1][2])) <= 2e-6 * scale


@pytest.mark.parametrize("case", [(4, 256, 38, 63, 1024, 1, True), (2, 128, 75, 125, 512, 1, True), (2, 256, 75, 125, 512, 2, False),
                                  (1, 192, 19, 32, 300, 1, False), (2, 64, 150, 250, 256, 1, True), (2, 64, 38, 63, 256, 1, False)])
def test_a_resident_variant_bit_identical(monkeypatch, case):
    """The A-resident sub-variant (the activation tile of an m tile converted into tensor memory once, the n tiles that
    follow stream weights only) issues the same MMAs on the same operands in the same order as the plain EPI2 kernel: the
    outputs are bit-identical, launch after launch."""
    N, Cin, H, W, Cout, stride, use_res = case
    g = torch.Generator(device="cuda").manual_seed(31 + Cin + Cout)
    x = torch.randn(N, Cin, H, W, device="cuda", generator=g)
    w = torch.randn(Cout, Cin, 1, 1, device="cuda", generator=g) * (2.0 / Cin) ** 0.5
    sc, sh = torch.rand(Cout, device="cuda", generator=g) + 0.5, torch.randn(Cout, device="cuda", generator=g)
    OH, OW = (H - 1) // stride + 1, (W - 1) // stride + 1
    res = dc.ActTensor.from_nchw(torch.randn(N, Cout, OH, OW, device="cuda", generator=g), cstride=Cout) if use_res else None
    outs = []
    for ares in ("0", "1"):
        monkeypatch.setenv("D2T_CONV_ARES", ares)
        layer = dc.ConvLayer(dc.ActTensor.from_nchw(x), w, sc, sh, stride, 0, 1, True, res, passes=16)
        for _ in range(3):
            layer.out.x.fill_(-3.0)
            layer.run()
        torch.cuda.synchronize()
        outs.append((layer.out.x.clone(), layer.out.amax.clone()))
    want = _ref(x, w, sc, sh, stride, 0, 1, True, res.to_nchw() if use_res else None)
    assert float((outs[1][0][..., :Cout].permute(0, 3, 1, 2) - want).abs().max() / want.abs().max()) < 1e-5
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


def test_two_engines_on_two_streams_with_private_scratch():
    """Two conv chains may overlap on different streams only with their own stream-K scratch (partial tiles + flags):
    results must equal the sequential runs (sharing the device-wide scratch would deadlock a finisher on a clobbered flag)."""
    from model.faster_rcnn.resnet import resnet
    from d2t_b200.engine import D2TEngine
    torch.manual_seed(3)
    net = resnet(tuple(range(31)), 50, class_agnostic=True).create_architecture().cuda().eval()
    H, W = 160, 224
    g = torch.Generator().manual_seed(2)
    ims = [(torch.rand(1, 2, 3, H, W, generator=g) * 256 - 128).cuda() for _ in range(2)]
    info = torch.tensor([H, W, 1.0]).view(1, 1, 3).expand(1, 2, 3).contiguous().cuda()
    engines = [D2TEngine(net, 1, H, W, private_scratch=True) for _ in range(2)]
    want = [[t.clone() for t in e(im, info)[:4]] for e, im in zip(engines, ims)]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for rep in range(3):
        outs = []
        for e, im, st in zip(engines, ims, streams):
            st.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(st):
                outs.append([t.clone() for t in e(im, info)[:4]])
        for st in streams:
            torch.cuda.current_stream().wait_stream(st)
        torch.cuda.synchronize()
        for o, w_ in zip(outs, want):
            for a, b in zip(o, w_):
                assert torch.equal(a, b)


def test_random_geometries_fp16_split_and_determinism():
    """3xFP16 mode on 24 random geometries / magnitudes (1e-3 .. 1e6) against float64, three runs each: <= 1e-5 of the
    output scale and bit-identical from run to run (stream-K fix-up order, TMA / TMEM hand-offs)."""
    rng = np.random.RandomState(5)
    for it in range(24):
        k = int(rng.choice([1, 1, 3]))
        stride = int(rng.choice([1, 1, 2])) if k == 1 else 1
        dil = int(rng.choice([1, 2, 6])) if k == 3 else 1
        pad = dil if k == 3 else 0
        N, Cin = int(rng.randint(1, 5)), int(rng.choice([32, 64, 96, 256, 320, 1024]))
        Cout = int(rng.choice([4, 24, 64, 128, 196, 260, 512, 1024]))
        H, W = int(rng.randint(3, 80)), int(rng.randint(3, 140))
        relu, use_res = bool(rng.randint(2)), bool(rng.randint(2)) and Cout % 4 == 0
        mag = float(10.0 ** rng.uniform(-3, 6))
        g = torch.Generator(device="cuda").manual_seed(2000 + it)
        x = torch.randn(N, Cin, H, W, device="cuda", generator=g) * mag
        w = torch.randn(Cout, Cin, k, k, device="cuda", generator=g) * (2.0 / (Cin * k * k)) ** 0.5
        sc = torch.rand(Cout, device="cuda", generator=g) + 0.5
        sh = torch.randn(Cout, device="cuda", generator=g) * mag
        OH = (H + 2 * pad - dil * (k - 1) - 1) // stride + 1
        OW = (W + 2 * pad - dil * (k - 1) - 1) // stride + 1
        res = torch.randn(N, Cout, OH, OW, device="cuda", generator=g) * mag if use_res else None
        rs = dc.ActTensor.from_nchw(res, cstride=Cout) if use_res else None
        layer = dc.ConvLayer(dc.ActTensor.from_nchw(x), w, sc, sh, stride, pad, dil, relu, rs, passes=16,
                             want_nhwc=(Cout % 4 == 0), want_nchw=True)
        outs = []
        for _ in range(3):
            layer.run()
            torch.cuda.synchronize()
            outs.append(layer.out_nchw.clone())
        want = _ref(x, w, sc, sh, stride, pad, dil, relu, res)
        err = float((outs[0] - want).abs().max() / want.abs().max())
        case = (N, Cin, H, W, Cout, k, stride, pad, dil, relu, use_res, mag)
        assert err < 1e-5, (case, err, layer.info)
        assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2]), case


def test_conv_chain_bit_identical():
    """ConvChain (one persistent launch over a list of layers, grid-wide barriers only between dependent layers) computes
    bit for bit what the same plans compute launched one by one: a bottleneck with a downsample branch (independent
    neighbours, EPI2 residual layer, BN = 64 and BN = 128 variants, stream-K splits), repeated launches included."""
    g = torch.Generator(device="cuda").manual_seed(11)

    def w(o, i, k):
        return torch.randn(o, i, k, k, device="cuda", generator=g) * (2.0 / (i * k * k)) ** 0.5

    def bn(c):
        return torch.rand(c, device="cuda", generator=g) + 0.5, torch.randn(c, device="cuda", generator=g) * 0.1

    for (N, H, W, cin, mid, cout) in [(2, 38, 63, 256, 64, 256), (4, 38, 63, 1024, 256, 1024), (1, 75, 125, 128, 128, 512)]:
        x = dc.ActTensor.from_nchw(torch.randn(N, cin, H, W, device="cuda", generator=g))

        def build():
            layers = []
            layers.append(dc.ConvLayer(x, w(cout, cin, 1), *bn(cout), passes=16))          # downsample-style branch
            res = layers[-1].out
            layers.append(dc.ConvLayer(x, w(mid, cin, 1), *bn(mid), relu=True, passes=16))
            layers.append(dc.ConvLayer(layers[-1].out, w(mid, mid, 3), *bn(mid), pad=2, dil=2, relu=True, passes=16))
            layers.append(dc.ConvLayer(layers[-1].out, w(cout, mid, 1), *bn(cout), relu=True, residual=res, passes=16))
            layers.append(dc.ConvLayer(layers[-1].out, w(mid, cout, 1), *bn(mid), relu=True, passes=16))
            layers.append(dc.ConvLayer(layers[-1].out, w(24, mid, 1), None, torch.randn(24, device="cuda", generator=g),
                                       passes=16, want_nhwc=False, want_nchw=True))
            return layers

        st = g.get_state()
        a = build()
        g.set_state(st)
        b = build()
        for l in a + b:                      # stand-alone layers zero their own amax; a chain leaves that to its owner
            l.zero_amax = None
        for l in a:
            l.run()
        chain = dc.ConvChain(b)
        assert chain.sync_before == [0, 0, 1, 1, 1, 1], chain.sync_before
        for rep in range(3):
            for l in b:
                if l.out is not None:
                    l.out.x.fill_(-7.0)
            chain.run()
            torch.cuda.synchronize()
            for la, lb in zip(a, b):
                if la.out is not None:
                    assert torch.equal(la.out.x, lb.out.x), (N, H, W, cin, rep)
                    assert torch.equal(la.out.amax, lb.out.amax)
                else:
                    assert torch.equal(la.out_nchw, lb.out_nchw)


def test_engine_chain_bit_identical():
    """D2TEngine with its layers collapsed into persistent chain launches == the layer-by-layer engine, bit for bit,
    eager and as a CUDA-graph replay."""
    from model.faster_rcnn.resnet import resnet
    from d2t_b200.engine import D2TEngine, GraphedEngine
    torch.manual_seed(3)
    net = resnet(tuple(range(31)), 50, class_agnostic=True).create_architecture().cuda().eval()
    B, H, W = 2, 224, 320
    gen = torch.Generator().manual_seed(1)
    im_data = (torch.rand(B, 2, 3, H, W, generator=gen) * 256 - 128).cuda()
    im_info = torch.tensor([H, W, 1.0]).view(1, 1, 3).expand(B, 2, 3).contiguous().cuda()
    plain = D2TEngine(net, B, H, W, chain=False)
    chained = D2TEngine(net, B, H, W, chain=True)
    assert len(chained.run_list) < len(plain.run_list) // 4
    want = [t.clone() for t in plain(im_data, im_info)[:4]]
    got = chained(im_data, im_info)[:4]
    torch.cuda.synchronize()
    for a, b in zip(want, got):
        assert torch.equal(a, b)
    assert torch.equal(plain.base_feat.x, chained.base_feat.x)
    graphed = GraphedEngine(chained, B, H, W)
    for _ in range(3):
        out = graphed(im_data, im_info)[:4]
        torch.cuda.synchronize()
        for a, b in zip(want, out):
            assert torch.equal(a, b)


def test_engine_notices_stale_parameters():
    """the engine computes with packed copies of the parameters: an in-place update after it was built must raise, not be
    silently ignored"""
    from model.faster_rcnn.resnet import resnet
    from d2t_b200.engine import D2TEngine
    torch.manual_seed(3)
    net = resnet(tuple(range(31)), 50, class_agnostic=True).create_architecture().cuda().eval()
    B, H, W = 1, 160, 224
    im_data = torch.zeros(B, 2, 3, H, W, device="cuda")
    im_info = torch.tensor([H, W, 1.0]).view(1, 1, 3).expand(B, 2, 3).contiguous().cuda()
    eng = D2TEngine(net, B, H, W)
    eng(im_data, im_info)
    net.load_state_dict(net.state_dict())          # (same values, but written in place)
    with pytest.raises(RuntimeError):
        eng(im_data, im_info)
    D2TEngine(net, B, H, W)(im_data, im_info)      # a fresh engine is fine
