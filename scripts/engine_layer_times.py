"""Per-layer device times of the engine's conv layers in place (warm L2 state of the real pipeline)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-detect-to-track_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from model.faster_rcnn.resnet import resnet
from d2t_b200.engine import D2TEngine
from d2t_b200 import conv as dc
torch.manual_seed(3)
net = resnet(tuple(range(31)), 101, class_agnostic=True).create_architecture().cuda().eval()
B, H, W = 2, 600, 1000
im_data = (torch.rand(B, 2, 3, H, W) * 256 - 128).cuda()
im_info = torch.tensor([H, W, 1.0]).view(1, 1, 3).expand(B, 2, 3).contiguous().cuda()
eng = D2TEngine(net, B, H, W, passes=int(sys.argv[1]) if len(sys.argv) > 1 else 3)
for _ in range(3):
    eng(im_data, im_info)
torch.cuda.synchronize()
frames = im_data.permute(1, 0, 2, 3, 4).reshape(2 * B, 3, H, W).contiguous()
reps = 5
acc = [0.0] * len(eng.layers)
for _ in range(reps):
    eng.stem.run(frames)
    dc.maxpool3x3s2(eng.stem.out, out=eng.pool_out)
    evs = []
    for layer in eng.layers:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); layer.run(); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    for i, (a, b) in enumerate(evs):
        acc[i] += a.elapsed_time(b) / reps
tot = 0.0
groups = {}
for i, (layer, ms) in enumerate(zip(eng.layers, acc)):
    x = layer.x
    O = layer.w_hi.shape[0]
    K = layer.w_hi.shape[1]
    key = "in[%dx%dx%d] K=%d -> %d %s%s" % (x.H, x.W, x.cstride, K, O, "res " if layer.residual is not None else "", "nchw" if layer.out_nchw is not None else "")
    g = groups.setdefault(key, [0, 0.0, layer.flops])
    g[0] += 1; g[1] += ms
    tot += ms
print("conv layers total %.3f ms (event-timed in place, incl. launch gaps)" % tot)
for k, (n, ms, fl) in sorted(groups.items(), key=lambda kv: -kv[1][1]):
    print("%-52s n=%3d  total %7.3f ms  each %6.1f us  %6.1f TF/s useful" % (k, n, ms, ms / n * 1e3, fl / (ms / n) / 1e9))
