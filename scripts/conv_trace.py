"""Per-role wait-cycle accounting of one conv layer (debug build: make -C pytorch-detect-to-track_b200/csrc trace).
usage: D2T_B200_LIB=.../libd2t_b200_trace.so python scripts/conv_trace.py N Cin H W Cout k stride pad dil passes [res]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-detect-to-track_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from d2t_b200 import conv as dc
from d2t_b200._lib import lib
N, Cin, H, W, Cout, k, stride, pad, dil, passes = [int(a) for a in sys.argv[1:11]]
x = torch.randn(N, Cin, H, W, device="cuda")
w = torch.randn(Cout, Cin, k, k, device="cuda") * 0.05
sc, sh = torch.rand(Cout, device="cuda") + 0.5, torch.randn(Cout, device="cuda")
OH = (H + 2 * pad - dil * (k - 1) - 1) // stride + 1
OW = (W + 2 * pad - dil * (k - 1) - 1) // stride + 1
res = dc.ActTensor.from_nchw(torch.randn(N, Cout, OH, OW, device="cuda"), cstride=Cout) if len(sys.argv) > 11 else None
layer = dc.ConvLayer(dc.ActTensor.from_nchw(x), w, sc, sh, stride, pad, dil, True, res, passes=passes)
trace = torch.zeros(148 * 8 * 8, dtype=torch.int64, device="cuda")
fn = lib().d2t_conv_plan_set_trace
fn.restype, fn.argtypes = C.c_int, [C.c_void_p, C.c_void_p]
grid = fn(layer.plan, trace.data_ptr())
flush = torch.zeros(64 * 1024 * 1024, device="cuda")
layer.zero_amax = None
for _ in range(3):
    flush.add_(1.0)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); layer.run(); layer.run(); layer.run(); b.record()      # (the trace is the last launch's)
    torch.cuda.synchronize()
print("layer", sys.argv[1:], "info", layer.info, "time %.1f us per launch (3 back to back)" % (a.elapsed_time(b) * 1e3 / 3))
t = trace.view(148, 8, 8)[:grid].double().cpu()
names = ["producer  [total, wait empty]", "mma       [total, wait xempty, wait tempty, wait full, wait cvt]",
         "cvt0      [total, wait full]", "cvt1", "cvt2", "cvt3", "epilogue0 [total, wait tfull, post-accumulate part, of which wait residual, post up to the end of the affine loop]", "epilogue1"]
print("kernel timeline (cycles from CTA entry): prologue done %d, previous grid complete %d, all roles done %d, TMEM released %d" %
      (t[:, 0, 5].mean(), t[:, 0, 6].mean(), t[:, 0, 7].mean(), t[:, 1, 5].mean()))
print("   epilogue warp 0 done at %d, final store wait %d cycles, dealloc %d cycles" % (t[:, 2, 6].mean(), t[:, 2, 5].mean(), t[:, 2, 7].mean()))
for r in range(8):
    m = t[:, r, :5]
    print("%-70s mean %s   max-total %d" % (names[r], [int(v) for v in m.mean(0)], int(m[:, 0].max())))
