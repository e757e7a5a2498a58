"""One conv layer, a few runs (for ncu).  usage: conv_prof1.py N Cin H W Cout k stride pad dil passes [res]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-detect-to-track_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from d2t_b200 import conv as dc
N, Cin, H, W, Cout, k, stride, pad, dil, passes = [int(a) for a in sys.argv[1:11]]
x = torch.randn(N, Cin, H, W, device="cuda")
w = torch.randn(Cout, Cin, k, k, device="cuda") * 0.05
sc, sh = torch.rand(Cout, device="cuda") + 0.5, torch.randn(Cout, device="cuda")
OH = (H + 2 * pad - dil * (k - 1) - 1) // stride + 1
OW = (W + 2 * pad - dil * (k - 1) - 1) // stride + 1
res = dc.ActTensor.from_nchw(torch.randn(N, Cout, OH, OW, device="cuda"), cstride=Cout) if len(sys.argv) > 11 else None
layer = dc.ConvLayer(dc.ActTensor.from_nchw(x), w, sc, sh, stride, pad, dil, True, res, passes=passes)
for _ in range(4):
    layer.run()
torch.cuda.synchronize()
