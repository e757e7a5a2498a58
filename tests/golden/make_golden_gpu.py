"""Runs THE REFERENCE'S OWN CUDA KERNELS (oracle/_ref/libref_oracle.so = the six unmodified
reference .cu files compiled for sm_100a by `make -C oracle ref`) on the seeded inputs of
cases.py and stores their outputs as tests/golden/ref_cuda.npz.  Needs a B200:

    gpurun -- 'python tests/golden/make_golden_gpu.py gpurun_out/golden'   # then copy the .npz here
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import common  # noqa: E402,F401
import cases  # noqa: E402
from oracle import ref_cuda as ref  # noqa: E402

dst = sys.argv[1] if len(sys.argv) > 1 else HERE
os.makedirs(dst, exist_ok=True)
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
out = {}

for name, c in cases.psroi_cases().items():
    top, mapping = ref.psroi_forward(cu(c["feat"]), cu(c["rois"]), c["scale"], c["P"], c["P"], c["G"], c["D"])
    g = ref.psroi_backward(cu(c["gtop"]), mapping, cu(c["rois"]), c["feat"].shape, c["scale"], c["P"], c["P"], c["D"])
    out["psroi_%s_top" % name], out["psroi_%s_map" % name] = top.cpu().numpy(), mapping.cpu().numpy()
    out["psroi_%s_grad" % name] = g.cpu().numpy()

for name, (dets, thresh) in cases.nms_cases().items():
    out["nms_%s" % name] = ref.nms(cu(dets), thresh).cpu().numpy().reshape(-1)

for name, c in cases.corr_cases().items():
    a, b, p = cu(c["in1"]), cu(c["in2"]), c["params"]
    o = ref.correlation_forward(a, b, *p)
    out["corr_%s_out" % name] = o.cpu().numpy()
    go = cu(common.randn(tuple(o.shape), c["gseed"]))
    slack = a.numel() * 4 if p[3] > 1 else 0   # stride1 > 1: the reference writes past the tensor
    g1, g2 = ref.correlation_backward(a, b, go, *p, slack=slack)
    out["corr_%s_g1" % name], out["corr_%s_g2" % name] = g1.cpu().numpy(), g2.cpu().numpy()

c = cases.roi_cases()
feat, rois, grid = cu(c["feat"]), cu(c["rois"]), cu(c["grid"])
for ah in (7, 8):
    top = ref.roi_align_forward(feat, rois, c["scale"], ah, ah)
    out["align%d_top" % ah] = top.cpu().numpy()
    gt = cu(common.randn(tuple(top.shape), 80 + ah))
    out["align%d_grad" % ah] = ref.roi_align_backward(gt, rois, c["feat"].shape, c["scale"], ah, ah).cpu().numpy()
top, arg = ref.roi_pool_forward(feat, rois, c["scale"], 7, 7)
out["pool_top"], out["pool_arg"] = top.cpu().numpy(), arg.cpu().numpy()
gt = cu(common.randn(tuple(top.shape), 90))
out["pool_grad"] = ref.roi_pool_backward(gt, arg, rois, c["feat"].shape, c["scale"], 7, 7).cpu().numpy()
o = ref.roi_crop_forward(feat, grid)
out["crop_out"] = o.cpu().numpy()
gi, gg = ref.roi_crop_backward(feat, grid, cu(common.randn(tuple(o.shape), 91)))
out["crop_gimg"], out["crop_ggrid_absmax"] = gi.cpu().numpy(), np.array(float(gg.abs().max()))
torch.cuda.synchronize()
np.savez_compressed(os.path.join(dst, "ref_cuda.npz"), **out)
print("wrote", os.path.join(dst, "ref_cuda.npz"), {k: v.shape for k, v in out.items()})
