"""Stress: random conv geometries in 3xFP16 mode against float64, and run-to-run bitwise determinism of the whole engine
(any race in the TMA / TMEM / stream-K hand-offs shows up as a differing bit)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-detect-to-track_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
import torch.nn.functional as F
from d2t_b200 import conv as dc
torch.backends.cudnn.allow_tf32 = False
rng = np.random.RandomState(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
worst = 0.0
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 60):
    k = int(rng.choice([1, 1, 3]))
    stride = int(rng.choice([1, 1, 2])) if k == 1 else 1
    dil = int(rng.choice([1, 2, 6])) if k == 3 else 1
    pad = dil if k == 3 else 0
    N = int(rng.randint(1, 5)); Cin = int(rng.choice([32, 64, 96, 128, 256, 320, 512, 1024]))
    Cout = int(rng.choice([4, 24, 64, 128, 196, 256, 260, 512, 1024]))
    H = int(rng.randint(3, 80)); W = int(rng.randint(3, 140))
    relu = bool(rng.randint(2)); use_res = bool(rng.randint(2)) and Cout % 4 == 0
    scale_mag = float(10.0 ** rng.uniform(-3, 6))
    g = torch.Generator(device="cuda").manual_seed(1000 + it)
    x = torch.randn(N, Cin, H, W, device="cuda", generator=g) * scale_mag
    w = torch.randn(Cout, Cin, k, k, device="cuda", generator=g) * (2.0 / (Cin * k * k)) ** 0.5
    sc = torch.rand(Cout, device="cuda", generator=g) + 0.5
    sh = torch.randn(Cout, device="cuda", generator=g) * scale_mag
    OH = (H + 2 * pad - dil * (k - 1) - 1) // stride + 1
    OW = (W + 2 * pad - dil * (k - 1) - 1) // stride + 1
    if OH <= 0 or OW <= 0:
        continue
    res = torch.randn(N, Cout, OH, OW, device="cuda", generator=g) * scale_mag if use_res else None
    rs = dc.ActTensor.from_nchw(res, cstride=Cout) if use_res else None
    layer = dc.ConvLayer(dc.ActTensor.from_nchw(x), w, sc, sh, stride, pad, dil, relu, rs, passes=16,
                         want_nhwc=(Cout % 4 == 0), want_nchw=True)
    outs = []
    for rep in range(3):
        layer.run()
        torch.cuda.synchronize()
        outs.append(layer.out_nchw.clone())
    y = F.conv2d(x.double(), w.double(), None, stride, pad, dil) * sc.double().view(1, -1, 1, 1) + sh.double().view(1, -1, 1, 1)
    if use_res:
        y = y + res.double()
    y = (F.relu(y) if relu else y).float()
    err = float((outs[0] - y).abs().max() / y.abs().max())
    worst = max(worst, err)
    same = all(torch.equal(outs[0], o) for o in outs[1:])
    if layer.out is not None:
        same = same and torch.equal(layer.out.to_nchw(Cout), outs[-1])
    info = (N, Cin, H, W, Cout, k, stride, pad, dil, relu, use_res, "%.1e" % scale_mag)
    if err >= 1e-5 or not same:
        print("FAIL", info, "err %.2e" % err, "deterministic", same, layer.info)
        sys.exit(1)
print("conv stress ok, worst rel err %.2e" % worst)
from model.faster_rcnn.resnet import resnet
from d2t_b200.engine import D2TEngine
torch.manual_seed(3)
net = resnet(tuple(range(31)), 101, class_agnostic=True).create_architecture().cuda().eval()
B, H, W = 2, 600, 1000
im = (torch.rand(B, 2, 3, H, W) * 256 - 128).cuda()
info = torch.tensor([H, W, 1.0]).view(1, 1, 3).expand(B, 2, 3).contiguous().cuda()
eng = D2TEngine(net, B, H, W)
ref = [t.clone() for t in eng(im, info)[:4]] + [eng.cls_map.clone(), eng.base_feat.x.clone()]
for rep in range(15):
    out = list(eng(im, info)[:4]) + [eng.cls_map, eng.base_feat.x]
    torch.cuda.synchronize()
    for a, b in zip(out, ref):
        if not torch.equal(a, b):
            print("engine run %d differs: max |d| %.3e" % (rep, float((a - b).abs().max())))
            sys.exit(1)
print("engine: 15 repeated forwards bit-identical")
