"""per-phase cycles of psroi_fwd_isat_mc (trace build: make -C csrc trace; D2T_B200_LIB=.../libd2t_b200_trace.so)"""
import sys, ctypes
sys.path.insert(0, '.'); sys.path.insert(0, 'pytorch-detect-to-track_b200')
import numpy as np
import torch
from d2t_b200 import ops, synth
from d2t_b200._lib import lib
B, D, R = 2, 30, 2000
rois = torch.from_numpy(synth.make_rois(R, B, seed=21)).cuda()
feat = torch.randn(B, D * 49, 38, 63, device='cuda')
for _ in range(3):
    ops.psroi_forward(feat, rois, 7, 7, 1 / 16., 7, D)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (480 * 8))()
lib().d2t_psroi_trace_read(buf)
a = np.array(buf).reshape(480, 8)[:420]
names = ["wait planes", "norm+scale", "quantise+row scan", "wait rows", "col scan", "wait cols", "lookups+stores"]
print("psroi_fwd_isat_mc, cycles per CTA (one item each, thread 0):", " | ".join("%s %d" % (n, a[:, i].mean()) for i, n in enumerate(names)), "| total %d" % a[:, :7].sum(1).mean())
