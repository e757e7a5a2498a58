"""PSRoI backward variants at BASELINE config 5 (D = 30, 2000 rois / image): time, HBM fraction, difference from the fp64 tables."""
import sys, json
sys.path.insert(0, '.')
import torch
sys.argv = ['bench.py']
import bench
from d2t_b200 import synth as common
from d2t_b200 import ops
from d2t_b200._lib import lib
torch.cuda.set_device(0)
flush = torch.zeros(128 * 1024 * 1024, device='cuda')
D = 30
for B in (1, 2, 8):
    R = 2000
    torch.manual_seed(20)
    rois = torch.from_numpy(common.make_rois(R, B, seed=21)).cuda()
    gt = torch.randn(B * R, D, 7, 7, device='cuda')
    shape = (B, D * 49, 38, 63)
    lib().d2t_psroi_set_mode(-1, 2)
    want = ops.psroi_backward(gt, rois, shape, 7, 7, 1 / 16., 7, D)
    alg = 4.0 * (D * 49 * 2394 + 5 * R + R * D * 49) * B
    for name, mode in (("fp64_cas", 2), ("limb", 0), ("int_mc", 1)):
        lib().d2t_psroi_set_mode(-1, mode)
        got = ops.psroi_backward(gt, rois, shape, 7, 7, 1 / 16., 7, D)
        ms = bench.time_kernel(lambda: ops.psroi_backward(gt, rois, shape, 7, 7, 1 / 16., 7, D), 20, flush)
        print(json.dumps({"mode": name, "B": B, "us": ms * 1e3, "gbs": alg / ms / 1e6, "frac_hbm": alg / ms / 1e6 / 6530.3,
                          "max_abs_diff_vs_fp64": float((got - want).abs().max()), "grad_absmax": float(want.abs().max()),
                          "nan": bool(torch.isnan(got).any())}), flush=True)
    lib().d2t_psroi_set_mode(-1, 0)
