"""Runs a handful of conv layers a few times each (for ncu).  usage: conv_prof.py passes"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-detect-to-track_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from d2t_b200 import conv as dc
passes = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for (N, Cin, H, W, Cout, k, stride, pad, dil) in [(4, 2048, 38, 63, 512, 3, 1, 6, 6), (4, 256, 38, 63, 1024, 1, 1, 0, 1),
                                                    (4, 256, 38, 63, 256, 3, 1, 1, 1), (4, 64, 150, 250, 256, 1, 1, 0, 1)]:
    x = torch.randn(N, Cin, H, W, device="cuda")
    w = torch.randn(Cout, Cin, k, k, device="cuda") * 0.05
    layer = dc.ConvLayer(dc.ActTensor.from_nchw(x), w, None, None, stride, pad, dil, True, None, passes=passes)
    for _ in range(3):
        layer.run()
    torch.cuda.synchronize()
