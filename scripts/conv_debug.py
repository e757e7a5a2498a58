"""Debug driver for the tcgen05 conv kernel: one case per process, prints an error map.
usage: python scripts/conv_debug.py CASE_INDEX PASSES"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-detect-to-track_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import torch.nn.functional as F
from d2t_b200 import conv as dc
from test_conv_gpu import CASES, _ref

idx, passes = int(sys.argv[1]), int(sys.argv[2])
N, Cin, H, W, Cout, k, stride, pad, dil, relu, use_res = CASES[idx]
g = torch.Generator(device="cuda").manual_seed(7)
x = torch.randn(N, Cin, H, W, device="cuda", generator=g)
w = torch.randn(Cout, Cin, k, k, device="cuda", generator=g) * (2.0 / (Cin * k * k)) ** 0.5
xs = dc.ActTensor.from_nchw(x)
print("layout round trip exact:", torch.equal(xs.to_nchw(Cin), x), flush=True)
layer = dc.ConvLayer(xs, w, None, None, stride, pad, dil, False, None, passes=passes, want_nhwc=(Cout % 4 == 0), want_nchw=True)
print("plan:", layer.info, flush=True)
layer.run()
torch.cuda.synchronize()
print("kernel finished", flush=True)
want = _ref(x, w, None, None, stride, pad, dil, False, None)
got = layer.out_nchw
err = (got - want).abs()
print("max rel err %.3e  (max |want| %.3f, max |got| %.3f)" % (float(err.max() / want.abs().max()), float(want.abs().max()), float(got.abs().max())))
bad = err > 1e-3 * want.abs().max()
print("bad fraction %.4f" % float(bad.float().mean()))
if bad.any():
    print("bad per channel (first 16):", bad.float().mean((0, 2, 3))[:16].tolist())
    print("bad per row (first 8):", bad.float().mean((0, 1, 3))[:8].tolist())
    print("bad per col (first 16):", bad.float().mean((0, 1, 2))[:16].tolist())
    print("got[0,0,0,:8]", got[0, 0, 0, :8].tolist())
    print("want[0,0,0,:8]", want[0, 0, 0, :8].tolist())
