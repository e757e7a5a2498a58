"""per-phase cycles of psroi_bwd_limb (trace build: make -C csrc trace; D2T_B200_LIB=.../libd2t_b200_trace.so)"""
import sys, ctypes
sys.path.insert(0, '.'); sys.path.insert(0, 'pytorch-detect-to-track_b200')
import numpy as np
import torch
from d2t_b200 import ops, synth
from d2t_b200._lib import lib
B, D, R = 2, 30, 2000
rois = torch.from_numpy(synth.make_rois(R, B, seed=21)).cuda()
gt = torch.randn(B * R, D, 7, 7, device='cuda')
for mode in (0,):
    lib().d2t_psroi_set_mode(-1, mode)
    for _ in range(3):
        ops.psroi_backward(gt, rois, (B, D * 49, 38, 63), 7, 7, 1 / 16., 7, D)
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * (480 * 8))()
    lib().d2t_psroi_trace_read(buf)
    a = np.array(buf).reshape(480, 8)[:148]
    print("mode", mode, "cycles per CTA (3 items; thread 0): zero+scale %d | corners %d | wait others %d | row scans %d | wait rows %d | col scans+write %d | total %d"
          % (tuple(a[:, i].mean() for i in (0, 1, 2, 3, 5, 4)) + (a[:, :7].sum(1).mean(),)))
