"""Experiment: one engine with 2 pairs per launch vs two engines with 1 pair each on two streams (do the
per-launch dependency bubbles of one chain hide behind the other chain's kernels?).  Each engine has its own
stream-K scratch (private_scratch=True); run under `timeout`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-detect-to-track_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from model.faster_rcnn.resnet import resnet
from d2t_b200.engine import D2TEngine
torch.manual_seed(3)
net = resnet(tuple(range(31)), 101, class_agnostic=True).create_architecture().cuda().eval()
H, W = 600, 1000
im = (torch.rand(2, 2, 3, H, W) * 256 - 128).cuda()
info = torch.tensor([H, W, 1.0]).view(1, 1, 3).expand(2, 2, 3).contiguous().cuda()
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
e2 = D2TEngine(net, 2, H, W)
print("one engine, 2 pairs/launch: %.3f ms" % timeit(lambda: e2(im, info)))
ea, eb = D2TEngine(net, 1, H, W, private_scratch=True), D2TEngine(net, 1, H, W, private_scratch=True)
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
ima, imb, infa, infb = im[:1].contiguous(), im[1:].contiguous(), info[:1].contiguous(), info[1:].contiguous()
def two():
    cur = torch.cuda.current_stream()
    sa.wait_stream(cur); sb.wait_stream(cur)
    with torch.cuda.stream(sa):
        ea(ima, infa)
    with torch.cuda.stream(sb):
        eb(imb, infb)
    cur.wait_stream(sa); cur.wait_stream(sb)
print("two engines, 1 pair each, two streams: %.3f ms" % timeit(two))
print("one engine, 1 pair alone: %.3f ms" % timeit(lambda: ea(ima, infa)))
ref = e2(im, info)
two(); torch.cuda.synchronize()
oa, ob = ea(ima, infa), eb(imb, infb)
print("same rois as the 2-pair engine:", bool(torch.equal(oa[0][:, 0], ref[0][:, 0])), bool(torch.equal(ob[0][:, 0], ref[0][:, 1])))
