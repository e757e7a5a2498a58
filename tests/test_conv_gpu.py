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
]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("passes", [3, 1])
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
    xs = dc.SplitTensor.from_nchw(x, lo=True)
    # the split is exact
    assert torch.equal(xs.to_nchw(Cin), x)
    rs = dc.SplitTensor.from_nchw(res, cstride=Cout) if use_res else None
    layer = dc.ConvLayer(xs, w, scale, shift, stride, pad, dil, relu, rs, passes=passes, want_nhwc=(Cout % 4 == 0),
                         want_nchw=True)
    out = layer.run()
    torch.cuda.synchronize()
    want = _ref(x, w, scale, shift, stride, pad, dil, relu, res)
    got = layer.out_nchw
    assert got.shape == want.shape
    err = float((got - want).abs().max() / want.abs().max())
    tol = 2e-6 if passes == 3 else 3e-3
    assert err < tol, (err, layer.info)
    if layer.out is not None:
        assert torch.equal(layer.out.to_nchw(Cout), got)      # NHWC split output carries the same values
        assert float((layer.out.hi.view(-1).view(torch.int32) & 0x1fff).abs().max()) == 0   # hi is a TF32 value


def test_maxpool_ceil_mode():
    x = torch.randn(2, 64, 37, 50, device="cuda")
    xs = dc.SplitTensor.from_nchw(x)
    out = dc.maxpool3x3s2(xs).to_nchw()
    want = F.max_pool2d(x, 3, 2, 0, ceil_mode=True)
    assert out.shape == want.shape and torch.equal(out, want)
    x = torch.randn(1, 64, 300, 500, device="cuda")
    out = dc.maxpool3x3s2(dc.SplitTensor.from_nchw(x)).to_nchw()
    want = F.max_pool2d(x, 3, 2, 0, ceil_mode=True)
    assert out.shape == want.shape == (1, 64, 150, 250) and torch.equal(out, want)
