"""A/B of the persistent layer chain (dc.ConvChain) against per-layer launches on a layer3-like stack of bottlenecks
(N = 4 frames of 38x63, 1024 -> 256 -> 256 -> 1024 + residual): us per block, bit-identity of the results."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "pytorch-detect-to-track_b200"))
import torch
from d2t_b200 import conv as dc
from d2t_b200._lib import lib

nblocks = int(sys.argv[1]) if len(sys.argv) > 1 else 8
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
    layers = []
    cur = x
    for b in range(nblocks):
        a = dc.ConvLayer(cur, w(256, 1024, 1), *bn(256), relu=True, passes=16)
        c2 = dc.ConvLayer(a.out, w(256, 256, 3), *bn(256), pad=1, relu=True, passes=16)
        c3 = dc.ConvLayer(c2.out, w(1024, 256, 1), *bn(1024), relu=True, residual=cur, passes=16)
        layers += [a, c2, c3]
        cur = c3.out
scratch = torch.zeros(lib().d2t_conv_scratch_bytes(), dtype=torch.uint8, device="cuda")
for l in layers:
    l.set_scratch(scratch)
chain = dc.ConvChain(layers)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")

def reset():
    arena.buf.zero_()
    x.amax.copy_(x0amax)

def run_plain():
    for l in layers:
        l.run()

def timeit(fn, reps=10):
    ts = []
    for _ in range(reps + 3):
        reset(); flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    return sum(ts[3:]) / reps * 1e3

reset(); run_plain(); torch.cuda.synchronize()
want = cur.x.clone()
reset(); chain.run(); torch.cuda.synchronize()
same = torch.equal(want, cur.x)
tp, tc = timeit(run_plain), timeit(chain.run)
print("blocks %d: per-layer launches %.1f us/block, chain %.1f us/block, identical %s, sync_before %s" % (nblocks, tp / nblocks, tc / nblocks, same, chain.sync_before[:6]))
