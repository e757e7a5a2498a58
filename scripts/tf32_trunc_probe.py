"""Does tcgen05 kind::tf32 ignore the low 13 mantissa bits of its fp32 operands (truncate)?
Run the 3-pass conv with hi := x (untruncated) and compare with hi := trunc(x)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-detect-to-track_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import torch.nn.functional as F
from d2t_b200 import conv as dc
torch.manual_seed(0)
N, Cin, H, W, Cout = 2, 256, 38, 63, 256
x = torch.randn(N, Cin, H, W, device="cuda")
w = torch.randn(Cout, Cin, 3, 3, device="cuda") * 0.02
want = F.conv2d(x.double(), w.double(), None, 1, 1, 1).float()
for mode in ("trunc", "full"):
    xs = dc.SplitTensor.from_nchw(x)
    layer = dc.ConvLayer(xs, w, None, None, 1, 1, 1, False, None, passes=3, want_nchw=True)
    if mode == "full":
        xs.hi.add_(xs.lo)          # hi := x exactly (lo unchanged)
        layer.w_hi.add_(layer.w_lo)
    layer.run()
    torch.cuda.synchronize()
    got = layer.out_nchw
    print(mode, "max rel err %.3e" % float((got - want).abs().max() / want.abs().max()))
    if mode == "trunc":
        base = got.clone()
    else:
        print("bit-identical to truncated-hi run:", bool(torch.equal(base, got)), " max diff %.3e" % float((base - got).abs().max()))
