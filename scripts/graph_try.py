"""Does replaying the whole eval forward as a CUDA graph shorten the step? (experiment)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-detect-to-track_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from model.faster_rcnn.resnet import resnet
from d2t_b200.engine import D2TEngine
torch.manual_seed(3)
net = resnet(tuple(range(31)), 101, class_agnostic=True).create_architecture().cuda().eval()
B, H, W = 2, 600, 1000
im = (torch.rand(B, 2, 3, H, W) * 256 - 128).cuda()
info = torch.tensor([H, W, 1.0]).view(1, 1, 3).expand(B, 2, 3).contiguous().cuda()
eng = D2TEngine(net, B, H, W)
flush = torch.zeros(128 * 1024 * 1024, device="cuda")
def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.add_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        tot += a.elapsed_time(b)
    return tot / n
ref = eng(im, info)
print("eager  %.3f ms" % timeit(lambda: eng(im, info)))
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(2):
        eng(im, info)
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g):
        out = eng(im, info)
    print("graph  %.3f ms" % timeit(g.replay))
    g.replay(); torch.cuda.synchronize()
    print("same rois:", bool(torch.equal(out[0], ref[0])), " max |d cls_prob| %.2e" % float((out[1] - ref[1]).abs().max()))
except Exception as e:
    print("capture failed:", repr(e)[:300])
