"""One PSRoI forward of BASELINE config 5 (B images x 2000 rois, D = 30) for an ncu capture:
  ncu --set full --import-source on --clock-control none -k regex:psroi_fwd -c 2 -o gpurun_out/psroi python scripts/psroi_one.py 8"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-detect-to-track_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import common
from d2t_b200._lib import lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
D, R = 30, 2000
torch.manual_seed(20)
feat = torch.randn(B, D * 49, 38, 63, device="cuda")
rois = torch.from_numpy(common.make_rois(R, B, seed=21)).cuda()
top = torch.empty(B * R, D, 7, 7, device="cuda")
ws = torch.empty(lib().d2t_psroi_workspace_bytes(B * R, B, 7, 7), dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    assert lib().d2t_psroi_forward(feat.data_ptr(), B, D * 49, 38, 63, rois.data_ptr(), B * R, 1 / 16., 7, 7, 7, D,
                                   top.data_ptr(), None, ws.data_ptr(), ws.numel(), st) == 1
torch.cuda.synchronize()
print("ok", float(top.abs().mean()))
