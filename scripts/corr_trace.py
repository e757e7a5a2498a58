"""Per-role wait-cycle accounting of the tensor-core correlation (debug build: make -C pytorch-detect-to-track_b200/csrc trace).
usage: D2T_B200_LIB=.../libd2t_b200_trace.so python scripts/corr_trace.py B C H W stride"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-detect-to-track_b200"))
import torch
from d2t_b200 import conv as dc
from d2t_b200._lib import lib
B, Cc, H, W, stride = [int(a) for a in sys.argv[1:6]]
a, b = torch.randn(B, Cc, H, W, device="cuda"), torch.randn(B, Cc, H, W, device="cuda")
layer = dc.CorrLayer(dc.ActTensor.from_nchw(a), dc.ActTensor.from_nchw(b), 8, 8, stride, passes=16, want_nchw=True)
layer.zero_amax = None
trace = torch.zeros(148 * 8 * 8, dtype=torch.int64, device="cuda")
fn = lib().d2t_conv_plan_set_trace
fn.restype, fn.argtypes = C.c_int, [C.c_void_p, C.c_void_p]
grid = fn(layer.plan, trace.data_ptr())
flush = torch.zeros(64 * 1024 * 1024, device="cuda")
for _ in range(3):
    flush.add_(1.0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); layer.run(); e1.record(); torch.cuda.synchronize()
print("corr", sys.argv[1:], "grid", grid, "time %.1f us (cold)" % (e0.elapsed_time(e1) * 1e3))
t = trace.view(148, 8, 8)[:grid].double().cpu()
names = ["producer [total, wait empty]", "mma [total, -, wait tempty, wait full, wait cvt]", "cvt0 [total, wait full]", "cvt1", "cvt2", "cvt3",
         "epilogue0 [total, wait tfull, post part, -, post up to the end of the band loop]", "epilogue1"]
print("timeline (cycles from CTA entry): prologue %d, previous grid complete %d, all roles done %d" % (t[:, 0, 5].mean(), t[:, 0, 6].mean(), t[:, 0, 7].mean()))
for r in range(8):
    m = t[:, r, :5]
    print("%-80s mean %s max-total %d" % (names[r], [int(v) for v in m.mean(0)], int(m[:, 0].max())))
