"""Back-to-back launches of one conv layer without events in between: average device time per launch
(fixed launch / prologue / tail cost shows up for tiny layers).  usage: conv_overhead.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-detect-to-track_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from d2t_b200 import conv as dc
from d2t_b200._lib import lib
from d2t_b200.conv import _stream
CASES = [(1, 64, 8, 16, 64, 1, 0), (4, 64, 38, 63, 64, 1, 0), (4, 512, 38, 63, 48, 1, 0), (4, 1024, 38, 63, 256, 1, 0),
         (4, 256, 38, 63, 1024, 1, 0), (4, 256, 38, 63, 256, 3, 1)]
for (N, Cin, H, W, Cout, k, pad) in CASES:
    for passes in (16, 3):
        x = torch.randn(N, Cin, H, W, device="cuda")
        w = torch.randn(Cout, Cin, k, k, device="cuda") * 0.05
        layer = dc.ConvLayer(dc.ActTensor.from_nchw(x), w, None, None, 1, pad, 1, True, None, passes=passes)
        layer.zero_amax = None
        run = lambda: lib().d2t_conv_plan_run(layer.plan, _stream())
        for _ in range(5):
            run()
        torch.cuda.synchronize()
        reps = 100
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            run()
        b.record()
        torch.cuda.synchronize()
        t_many = a.elapsed_time(b) * 1e3 / reps
        a.record(); run(); b.record(); torch.cuda.synchronize()
        t_one = a.elapsed_time(b) * 1e3
        print("N%d Cin%d %dx%d Cout%d k%d passes %2d: %.1f us/launch back-to-back (x%d), %.1f us single, grid %d" %
              (N, Cin, H, W, Cout, k, passes, t_many, reps, t_one, layer.info["grid"] // 10))
