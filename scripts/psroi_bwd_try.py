"""PSRoI backward: the fp64 difference-table kernel (default) against the integer-table multi-CTA experiment
(D2T_PSROI_BWD_INT=1, csrc/psroi.cu psroi_bwd_isat_mc -- written at the end of round 1, NOT yet run on a GPU).
One JSON line per (mode, batch): time, GB/s against the algorithmic bytes, max |difference| to the fp64 kernel.
Run under `timeout`."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-detect-to-track_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
sys.argv = ["bench.py"]
import bench
import common
from d2t_b200._lib import lib

torch.cuda.set_device(0)
flush = torch.zeros(128 * 1024 * 1024, device="cuda")
HBM = bench.peaks()[0]
D, R = 30, 2000
for B in (1, 2, 8):
    torch.manual_seed(20)
    rois = torch.from_numpy(common.make_rois(R, B, seed=21)).cuda()
    gt = torch.randn(B * R, D, 7, 7, device="cuda")
    grad = torch.empty(B, D * 49, 38, 63, device="cuda")
    ws = torch.empty(lib().d2t_psroi_workspace_bytes(B * R, B, 7, 7), dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    alg = 4.0 * (D * 49 * 2394 + 5 * R + R * D * 49) * B
    ref = None
    for name, env in (("fp64", None), ("int_mc", "1")):
        if env is None:
            os.environ.pop("D2T_PSROI_BWD_INT", None)
        else:
            os.environ["D2T_PSROI_BWD_INT"] = env
        grad.fill_(float("nan"))

        def call():
            assert lib().d2t_psroi_backward(gt.data_ptr(), B, D * 49, 38, 63, rois.data_ptr(), B * R, 1 / 16., 7, 7, 7, D,
                                            grad.data_ptr(), 0, ws.data_ptr(), ws.numel(), st) == 1
        ms = bench.time_kernel(call, 10, flush)
        torch.cuda.synchronize()
        if ref is None:
            ref = grad.clone()
        print(json.dumps({"mode": name, "B": B, "us": ms * 1e3, "gbs": alg / ms / 1e6, "frac_hbm": alg / ms / 1e6 / HBM,
                          "max_abs_diff_vs_fp64": float((grad - ref).abs().max()), "grad_absmax": float(ref.abs().max()),
                          "nan": bool(torch.isnan(grad).any())}), flush=True)
os.environ.pop("D2T_PSROI_BWD_INT", None)
