"""Per-layer timing of the tcgen05 conv kernel on the Res-101 D&T shapes (4 frames of 600x1000),
next to cuDNN fp32 / TF32 through torch.  usage: python scripts/conv_bench.py [passes]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-detect-to-track_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import torch.nn.functional as F
from d2t_b200 import conv as dc

SHAPES = [  # name, N, Cin, H, W, Cout, k, stride, pad, dil, count per frame-batch
    ("l1 1x1 64->64", 4, 64, 150, 250, 64, 1, 1, 0, 1, 1),
    ("l1 3x3 64->64", 4, 64, 150, 250, 64, 3, 1, 1, 1, 3),
    ("l1 1x1 64->256", 4, 64, 150, 250, 256, 1, 1, 0, 1, 4),
    ("l1 1x1 256->64", 4, 256, 150, 250, 64, 1, 1, 0, 1, 2),
    ("l2 1x1 256->128 s2", 4, 256, 150, 250, 128, 1, 2, 0, 1, 1),
    ("l2 3x3 128->128", 4, 128, 75, 125, 128, 3, 1, 1, 1, 4),
    ("l2 1x1 128->512", 4, 128, 75, 125, 512, 1, 1, 0, 1, 4),
    ("l2 1x1 512->128", 4, 512, 75, 125, 128, 1, 1, 0, 1, 3),
    ("l3 3x3 256->256", 4, 256, 38, 63, 256, 3, 1, 1, 1, 23),
    ("l3 1x1 256->1024", 4, 256, 38, 63, 1024, 1, 1, 0, 1, 23),
    ("l3 1x1 1024->256", 4, 1024, 38, 63, 256, 1, 1, 0, 1, 22),
    ("l4 3x3 512->512 d2", 4, 512, 38, 63, 512, 3, 1, 2, 2, 3),
    ("l4 1x1 512->2048", 4, 512, 38, 63, 2048, 1, 1, 0, 1, 3),
    ("l4 1x1 2048->512", 4, 2048, 38, 63, 512, 1, 1, 0, 1, 2),
    ("l4 1x1 1024->2048", 4, 1024, 38, 63, 2048, 1, 1, 0, 1, 1),
    ("head 3x3 2048->512 d6", 4, 2048, 38, 63, 512, 3, 1, 6, 6, 1),
    ("rpn 3x3 512->512", 4, 512, 38, 63, 512, 3, 1, 1, 1, 1),
    ("cls 1x1 512->1519", 4, 512, 38, 63, 1519, 1, 1, 0, 1, 1),
]
passes = int(sys.argv[1]) if len(sys.argv) > 1 else 3
flush = torch.zeros(64 * 1024 * 1024, device="cuda")


def timeit(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.add_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        tot += a.elapsed_time(b)
    return tot / iters


rows, tot_mine, tot_cudnn, tot_tf32 = [], 0.0, 0.0, 0.0
for name, N, Cin, H, W, Cout, k, stride, pad, dil, cnt in SHAPES:
    x = torch.randn(N, Cin, H, W, device="cuda")
    w = torch.randn(Cout, Cin, k, k, device="cuda") * 0.05
    xs = dc.ActTensor.from_nchw(x)
    layer = dc.ConvLayer(xs, w, None, None, stride, pad, dil, True, None, passes=passes, want_nhwc=(Cout % 4 == 0), want_nchw=(Cout % 4 != 0))
    ms = timeit(layer.run)
    xl = x.contiguous(memory_format=torch.channels_last)
    torch.backends.cudnn.allow_tf32 = False
    ms_c = timeit(lambda: F.conv2d(x, w, None, stride, pad, dil))
    torch.backends.cudnn.allow_tf32 = True
    ms_t = timeit(lambda: F.conv2d(xl, w, None, stride, pad, dil))
    tf = layer.flops / ms / 1e9
    rows.append((name, layer.info["m_tiles"] * layer.info["n_tiles"], ms, tf, ms_c, ms_t))
    tot_mine += ms * cnt; tot_cudnn += ms_c * cnt; tot_tf32 += ms_t * cnt
    print("%-24s tiles %5d  d2t %7.3f ms %7.1f TF/s(useful) | cudnn fp32 %7.3f ms | cudnn tf32 %7.3f ms" % rows[-1], flush=True)
print("weighted trunk+heads total (4 frames): d2t(passes=%d) %.2f ms | cudnn fp32 %.2f ms | cudnn tf32 %.2f ms" % (passes, tot_mine, tot_cudnn, tot_tf32))
