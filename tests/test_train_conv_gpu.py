"""Backward-data (DgradConv) and weight-gradient (WgradLayer) of the tcgen05 convolution engine against
torch.autograd.grad of float64 conv2d on the same inputs -- the 3xFP16 bar of the forward pass: <= 1e-5 of the
result's scale.  Replaces the cuDNN dgrad / wgrad calls behind the reference's loss.backward()
(/root/reference/trainval_net.py:371-373)."""
import pytest
import torch
import torch.nn.functional as F

import common  # noqa: F401
from d2t_b200 import conv as dc

pytestmark = pytest.mark.gpu

CASES = [
    # N, Cin, H, W, Cout, k, stride, pad, dil
    (1, 64, 8, 16, 64, 1, 1, 0, 1),
    (2, 256, 38, 63, 64, 1, 1, 0, 1),          # 1x1, W = 63
    (2, 64, 38, 63, 256, 1, 1, 0, 1),          # expansion 1x1
    (2, 128, 38, 63, 128, 3, 1, 1, 1),         # 3x3 pad 1
    (1, 512, 38, 63, 512, 3, 1, 2, 2),         # layer4 dilated 3x3
    (1, 256, 19, 32, 512, 3, 1, 6, 6),         # dilation-6 head conv
    (2, 256, 75, 125, 128, 1, 2, 0, 1),        # stride-2 1x1 (first conv of a stage), W = 125 -> 63
    (1, 512, 38, 63, 1519, 1, 1, 0, 1),        # R-FCN cls head: Cout not a multiple of 16
    (1, 1051, 19, 32, 196, 1, 1, 0, 1),        # tracking head: Cin padded to 1056
    (2, 512, 38, 63, 24, 1, 1, 0, 1),          # RPN scores: BN = 64 tile, Cout padded to 64 as the dgrad's K
    (1, 128, 75, 125, 128, 3, 1, 1, 1),        # layer2 3x3, two K blocks per output row
]


def _setup(case, seed=0):
    N, Cin, H, W, Cout, k, stride, pad, dil = case
    g = torch.Generator(device="cuda").manual_seed(4321 + seed + Cin + Cout + k)
    x = torch.randn(N, Cin, H, W, device="cuda", generator=g)
    x = F.relu(x)                                                        # a post-ReLU activation: its zeros are the mask
    w = torch.randn(Cout, Cin, k, k, device="cuda", generator=g) * (2.0 / (Cin * k * k)) ** 0.5
    scale = torch.rand(Cout, device="cuda", generator=g) + 0.5
    OH = (H + 2 * pad - dil * (k - 1) - 1) // stride + 1
    OW = (W + 2 * pad - dil * (k - 1) - 1) // stride + 1
    gy = torch.randn(N, Cout, OH, OW, device="cuda", generator=g) * 1e-3   # gradient w.r.t. the conv's pre-activation output
    return x, w, scale, gy


def _ref_grads(x, w, scale, gy, stride, pad, dil):
    xd = x.double().requires_grad_()
    wd = w.double().requires_grad_()
    y = F.conv2d(xd, wd, None, stride, pad, dil) * scale.double().view(1, -1, 1, 1)
    gx, gw = torch.autograd.grad(y, (xd, wd), gy.double())
    return gx, gw


def _amax_of(t):
    return t.abs().max().reshape(1).float()


@pytest.mark.parametrize("case", CASES)
def test_wgrad_matches_autograd(case):
    N, Cin, H, W, Cout, k, stride, pad, dil = case
    x, w, scale, gy = _setup(case)
    _, gw_ref = _ref_grads(x, w, scale, gy, stride, pad, dil)
    xs = dc.ActTensor.from_nchw(x)
    gs = dc.ActTensor.from_nchw(gy)
    nx, ng = dc.WgradScratch.need(xs, gs, stride, k)
    scratch = dc.WgradScratch(nx, ng)
    gw = torch.full_like(w, float("nan"))
    layer = dc.WgradLayer(xs, gs, gw, scale, stride, pad, dil, scratch)
    outs = []
    for _ in range(2):
        gw.fill_(float("nan"))
        layer.run()
        torch.cuda.synchronize()
        outs.append(gw.clone())
    assert bool(torch.isfinite(gw).all())
    err = float((gw.double() - gw_ref).abs().max() / gw_ref.abs().max())
    print("wgrad", case, "max rel err %.2e" % err)
    assert err < 1e-5, (case, err)
    assert torch.equal(outs[0], outs[1])                                  # deterministic (stream-K fix-up order)


@pytest.mark.parametrize("case", CASES)
def test_dgrad_matches_autograd(case):
    N, Cin, H, W, Cout, k, stride, pad, dil = case
    if stride != 1:
        pytest.skip("strided backward-data runs at the output resolution + scatter: test_dgrad_stride2_scatter")
    x, w, scale, gy = _setup(case)
    gx_ref, _ = _ref_grads(x, w, scale, gy, stride, pad, dil)
    g = torch.Generator(device="cuda").manual_seed(7)
    skip = torch.randn(N, Cin, H, W, device="cuda", generator=g) * float(gx_ref.abs().max())   # the skip connection's gradient
    want = (gx_ref + skip.double()) * (x > 0).double()
    gs = dc.ActTensor.from_nchw(gy)
    cpad = (Cin + 31) // 32 * 32 if Cin % 4 else Cin
    xs = dc.ActTensor.from_nchw(x, cstride=cpad)
    rs = dc.ActTensor.from_nchw(skip, cstride=cpad)
    out = dc.ActTensor(N, H, W, cpad, cstride=cpad)
    amax_wt = _amax_of(w * scale.view(-1, 1, 1, 1))
    layer = dc.DgradConv(gs, w, scale, pad, dil, amax_wt, out=out, residual=rs, mask=xs, out_channels=cpad)
    layer.run()
    torch.cuda.synchronize()
    got = out.to_nchw(Cin)
    err = float((got.double() - want).abs().max() / want.abs().max())
    print("dgrad", case, "max rel err %.2e" % err, layer.info)
    assert err < 1e-5, (case, err)
    assert float(out.amax) == float(got.abs().max())                      # the epilogue's running max |g|
    # in-place accumulation (residual == out), as the engine chains several consumers of one tensor
    out2 = dc.ActTensor(N, H, W, cpad, cstride=cpad)
    out2.x.copy_(rs.x)
    layer2 = dc.DgradConv(gs, w, scale, pad, dil, amax_wt, out=out2, residual=out2, mask=xs, out_channels=cpad)
    layer2.run()
    torch.cuda.synchronize()
    assert torch.equal(out2.x, out.x)


def test_dgrad_stride2_scatter():
    """first block of a stage: conv1 and the downsample conv are 1x1 stride 2 -- their backward-data runs at the output
    resolution (the second accumulating in place), then one kernel scatters to the even positions, adds the gradient
    arriving from elsewhere (the correlation) and applies the ReLU mask"""
    N, Cin, H, W = 2, 512, 75, 125
    g = torch.Generator(device="cuda").manual_seed(3)
    x = F.relu(torch.randn(N, Cin, H, W, device="cuda", generator=g))
    w1 = torch.randn(256, Cin, 1, 1, device="cuda", generator=g) * 0.05
    w2 = torch.randn(1024, Cin, 1, 1, device="cuda", generator=g) * 0.05
    s1, s2 = torch.rand(256, device="cuda", generator=g) + 0.5, torch.rand(1024, device="cuda", generator=g) + 0.5
    g1 = torch.randn(N, 256, 38, 63, device="cuda", generator=g)
    g2 = torch.randn(N, 1024, 38, 63, device="cuda", generator=g)
    extra = torch.randn(N, Cin, H, W, device="cuda", generator=g)
    xd = x.double().requires_grad_()
    y1 = F.conv2d(xd, w1.double(), None, 2) * s1.double().view(1, -1, 1, 1)
    y2 = F.conv2d(xd, w2.double(), None, 2) * s2.double().view(1, -1, 1, 1)
    (gx_ref,) = torch.autograd.grad([y1, y2], xd, [g1.double(), g2.double()])
    want = (gx_ref + extra.double()) * (x > 0).double()
    xs, es = dc.ActTensor.from_nchw(x, cstride=Cin), dc.ActTensor.from_nchw(extra, cstride=Cin)
    low = dc.ActTensor(N, 38, 63, Cin, cstride=Cin)
    out = dc.ActTensor(N, H, W, Cin, cstride=Cin)
    a = dc.DgradConv(dc.ActTensor.from_nchw(g1), w1, s1, 0, 1, _amax_of(w1 * s1.view(-1, 1, 1, 1)), out=low)
    b = dc.DgradConv(dc.ActTensor.from_nchw(g2), w2, s2, 0, 1, _amax_of(w2 * s2.view(-1, 1, 1, 1)), out=low, residual=low)
    a.run(), b.run()
    dc.upsample2_add_mask(low, out, extra=es, mask=xs)
    torch.cuda.synchronize()
    got = out.to_nchw(Cin)
    err = float((got.double() - want).abs().max() / want.abs().max())
    assert err < 1e-5, err
    assert float(out.amax) == float(got.abs().max())


def test_dynamic_weight_scale_repack():
    """forward plan with the weight scale on the device: after the weights change (an optimizer step) repack() alone
    brings the plan up to date -- no host synchronisation, no new plan"""
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(2, 128, 20, 30, device="cuda", generator=g)
    w = torch.nn.Parameter(torch.randn(96, 128, 3, 3, device="cuda", generator=g) * 0.05)
    amax = torch.zeros(1, device="cuda")
    amax.copy_(w.detach().abs().max())
    layer = dc.ConvLayer(dc.ActTensor.from_nchw(x), w, None, None, 1, 1, 1, False, None, passes=16, want_nchw=True, amax_w=amax)
    for scale in (1.0, 37.0, 1e-4):
        with torch.no_grad():
            w.mul_(scale)
        amax.copy_(w.detach().abs().max())
        layer.repack()
        layer.run()
        torch.cuda.synchronize()
        want = F.conv2d(x.double(), w.detach().double(), None, 1, 1, 1)
        err = float((layer.out_nchw.double() - want).abs().max() / want.abs().max())
        assert err < 1e-5, (scale, err)


@pytest.mark.parametrize("C,H,W,md,stride,B", [(128, 38, 63, 8, 1, 2), (1024, 38, 63, 8, 1, 1), (512, 75, 125, 8, 2, 2),
                                              (64, 20, 30, 8, 1, 1), (96, 21, 45, 4, 1, 1)])
def test_correlation_backward_tensor_core(C, H, W, md, stride, B):
    """CORRB mode of the tcgen05 kernel (banded GEMM over the halo positions) against the exact-adjoint fp32 SIMT kernels of
    csrc/correlation.cu, both gradients, every D&T geometry; for stride 2 the gradient lives on the even positions"""
    from d2t_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(55 + C)
    in1 = torch.randn(B, C, H, W, device="cuda", generator=g)
    in2 = torch.randn(B, C, H, W, device="cuda", generator=g)
    r = md // stride
    D2 = (2 * r + 1) ** 2
    oh, ow = -(-H // stride), -(-W // stride)
    go = torch.randn(B, D2, oh, ow, device="cuda", generator=g) * 1e-3
    want1, want2 = ops.correlation_backward(in1, in2, go, md, 1, md, stride, stride)
    coff = 12
    gbuf = dc.ActTensor(B, oh, ow, coff + D2 + 5, cstride=(coff + D2 + 5 + 3) // 4 * 4)
    gbuf.load_nchw(torch.cat([torch.randn(B, coff, oh, ow, device="cuda", generator=g), go,
                              torch.randn(B, 5, oh, ow, device="cuda", generator=g)], 1))
    x1, x2 = dc.ActTensor.from_nchw(in1), dc.ActTensor.from_nchw(in2)
    nb, no = dc.CorrBwdScratch.need(B, C, oh, ow, r)
    scratch = dc.CorrBwdScratch(nb, no)
    for which, other, want in ((1, x2, want1), (2, x1, want2)):
        out = dc.ActTensor(B, oh, ow, C, cstride=C)
        out.x.fill_(float("nan"))
        layer = dc.CorrBwdLayer(gbuf, coff, other, out, md, stride, which, scratch)
        layer.run()
        torch.cuda.synchronize()
        got = out.to_nchw(C)
        ref = want[:, :, ::stride, ::stride]
        assert got.shape == ref.shape
        err = float((got - ref).abs().max() / ref.abs().max())
        print("corr bwd", (C, H, W, md, stride, B), "d/d(input%d)" % which, "max rel err %.2e" % err)
        assert err < 2e-5, (which, err)
        if stride > 1:      # nothing off the lattice
            mask = torch.ones_like(want)
            mask[:, :, ::stride, ::stride] = 0
            assert float((want * mask).abs().max()) == 0.0
