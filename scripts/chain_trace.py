"""Per-layer timeline of the persistent layer chain (debug build: make -C pytorch-detect-to-track_b200/csrc trace).
usage: D2T_B200_LIB=.../libd2t_b200_trace.so python scripts/chain_trace.py [blocks]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-detect-to-track_b200"))
import torch
from d2t_b200 import conv as dc
from d2t_b200._lib import lib
nblocks = int(sys.argv[1]) if len(sys.argv) > 1 else 3
g = torch.Generator(device="cuda").manual_seed(1)
def w(o, i, k):
    return torch.randn(o, i, k, k, device="cuda", generator=g) * (2.0 / (i * k * k)) ** 0.5
def bn(c):
    return torch.rand(c, device="cuda", generator=g) * 0.2 + 0.4, torch.randn(c, device="cuda", generator=g) * 0.1
N, H, W = 4, 38, 63
arena = dc.AmaxArena(256)
with arena:
    x = dc.ActTensor.from_nchw(torch.randn(N, 1024, H, W, device="cuda", generator=g))
    x0amax = x.amax.clone()
    layers, cur = [], x
    for b in range(nblocks):
        a = dc.ConvLayer(cur, w(256, 1024, 1), *bn(256), relu=True, passes=16)
        c2 = dc.ConvLayer(a.out, w(256, 256, 3), *bn(256), pad=1, relu=True, passes=16)
        c3 = dc.ConvLayer(c2.out, w(1024, 256, 1), *bn(1024), relu=True, residual=cur, passes=16)
        layers += [a, c2, c3]
        cur = c3.out
scratch = torch.zeros(lib().d2t_conv_scratch_bytes(), dtype=torch.uint8, device="cuda")
fn = lib().d2t_conv_plan_set_trace
fn.restype, fn.argtypes = C.c_int, [C.c_void_p, C.c_void_p]
traces = []
for l in layers:
    l.set_scratch(scratch)
    t = torch.zeros(148 * 8 * 8, dtype=torch.int64, device="cuda")
    fn(l.plan, t.data_ptr())
    traces.append(t)
chain = dc.ConvChain(layers)
flush = torch.zeros(64 * 1024 * 1024, device="cuda")
for _ in range(3):
    arena.buf.zero_(); x.amax.copy_(x0amax); flush.add_(1.0)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); chain.run(); b.record(); torch.cuda.synchronize()
print("chain of %d layers: %.1f us (%.1f us / layer)" % (len(layers), a.elapsed_time(b) * 1e3, a.elapsed_time(b) * 1e3 / len(layers)))
T = [t.view(148, 8, 8).double().cpu() for t in traces]
names = ["producer[total, wait empty]", "mma[total, -, wait tempty, wait full, wait cvt]", "cvt0[total, wait full]", "epi0[total, wait tfull, post, wait res, affine]"]
for i, t in enumerate(T):
    left, passed, start, first, epi = t[:, 0, 6], t[:, 0, 7], t[:, 0, 5], t[:, 1, 5], t[:, 6, 5]
    nxt = T[i + 1][:, 0, 6] if i + 1 < len(T) else None
    line = "layer %2d (%s): " % (i, "abc"[i % 3])
    if i > 0:
        line += "barrier wait %6d | reset %5d | " % ((passed - left).mean(), (start - passed).mean())
    line += "start->first MMA %6d | first MMA->epilogue done %6d (max %6d)" % ((first - start)[first > 0].mean(), (epi - first)[first > 0].mean(), (epi - first)[first > 0].max())
    if nxt is not None:
        line += " | epilogue done->layer left %5d" % (nxt - epi).mean()
    print(line)
    for r, nm in zip((0, 1, 2, 6), names):
        m = t[:, r, :5]
        print("      %-48s mean %s max-total %d" % (nm, [int(v) for v in m.mean(0)], int(m[:, 0].max())))

# distribution of one (a) layer's main phase over the CTAs: who is the straggler?
import numpy as np
t = T[3]
dur = (t[:, 6, 5] - t[:, 1, 5]).numpy()
order = np.argsort(dur)
print("layer 3 (a): main phase per CTA, sorted: min %d p25 %d median %d p75 %d p90 %d max %d" % tuple(np.percentile(dur, [0, 25, 50, 75, 90, 100])))
print("slowest 12 CTAs (index: cycles, epilogue wait-tfull, post):", [(int(i), int(dur[i]), int(t[i, 6, 1]), int(t[i, 6, 2])) for i in order[-12:]])
print("fastest 6 CTAs:", [(int(i), int(dur[i]), int(t[i, 6, 1]), int(t[i, 6, 2])) for i in order[:6]])
mma = t[:, 1, :5].numpy()
print("mma totals of the slowest 6:", [(int(i), [int(v) for v in mma[i]]) for i in order[-6:]])
print("mma totals of the median 3:", [(int(i), [int(v) for v in mma[i]]) for i in order[72:75]])
