import os, sys, json
sys.path.insert(0, '.'); 
import torch
sys.argv=['bench.py']
import bench
torch.cuda.set_device(0)
flush = torch.zeros(128*1024*1024, device='cuda')
import common
from d2t_b200._lib import lib
B, D, R = 2, 30, 2000
torch.manual_seed(20)
feat = torch.randn(B, D*49, 38, 63, device='cuda')
rois = torch.from_numpy(common.make_rois(R, B, seed=21)).cuda()
top = torch.empty(B*R, D, 7, 7, device='cuda')
ws = torch.empty(lib().d2t_psroi_workspace_bytes(B*R, B, 7, 7), dtype=torch.uint8, device='cuda')
st = torch.cuda.current_stream().cuda_stream
def psroi():
    lib().d2t_psroi_forward(feat.data_ptr(), B, D*49, 38, 63, rois.data_ptr(), B*R, 1/16., 7, 7, 7, D, top.data_ptr(), None, ws.data_ptr(), ws.numel(), st)
ms = bench.time_kernel(psroi, 30, flush)
alg = 4.0*(D*49*2394 + 5*R + R*D*49)*B
print("psroi fwd %.1f us  %.1f GB/s  frac %.3f  (int-table experiment: %s)" % (ms*1e3, alg/ms/1e6, alg/ms/1e6/6530.3, os.environ.get("D2T_PSROI_INT")))
t1 = top.clone()
os.environ["D2T_PSROI_INT"] = "0"
psroi(); torch.cuda.synchronize()
print("max |int - fp64| = %.3e" % float((t1-top).abs().max()))

if hasattr(lib(), "d2t_psroi_trace_read"):
    import ctypes, numpy as np
    os.environ["D2T_PSROI_INT"] = "1"
    psroi(); torch.cuda.synchronize()
    buf = (ctypes.c_longlong * (480 * 8))()
    lib().d2t_psroi_trace_read(buf)
    a = np.array(buf).reshape(480, 8)[:148]
    print("per-CTA cycles (thread 0): wait data %d | pass 1 (load, L1) %d | scale %d... row scan %d | column scan %d | wait prep %d | lookups %d"
          % tuple(a[:, i].mean() for i in (0, 1, 1, 2, 3, 4, 5)))
    print("   total", a[:, :6].sum(1).mean())
