"""BASELINE config 4: correlation microbench, 1024x38x63 feature pairs, max_displacement 8, batch sweep 1-64 (plus the
conv3 / conv5 shapes at B = 2 and 16).  Tensor-core CORR mode of the conv kernel on the engine's NHWC layout; one JSON
line per point: time, GB/s against the algorithmic bytes (4 (2 C H W + D^2 oh ow) per pair), useful TFLOP/s.
Run under `timeout`."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-detect-to-track_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
sys.argv = ["bench.py"]
import bench
from d2t_b200 import conv as dc, ops

torch.cuda.set_device(0)
flush = torch.zeros(128 * 1024 * 1024, device="cuda")
HBM = bench.peaks()[0]
SHAPES = {"conv4": (1024, 38, 63, (8, 1, 8, 1, 1)), "conv5": (2048, 38, 63, (8, 1, 8, 1, 1)), "conv3": (512, 75, 125, (8, 1, 8, 2, 2))}


def point(name, B):
    C, H, W, p = SHAPES[name]
    a, b = torch.randn(B, C, H, W, device="cuda"), torch.randn(B, C, H, W, device="cuda")
    oc, oh, ow = ops.correlation_shape(H, W, *p)
    layer = dc.CorrLayer(dc.ActTensor.from_nchw(a), dc.ActTensor.from_nchw(b), p[0], p[2], p[3], passes=3, want_nchw=True)
    del a, b
    ms = bench.time_kernel(layer.run, 10, flush)
    touched = (oh * ow) if p[3] > 1 else H * W
    alg = 4.0 * (2 * C * touched + oc * oh * ow) * B
    flops = 2.0 * oc * oh * ow * C * B
    print(json.dumps({"shape": name, "B": B, "us": ms * 1e3, "us_per_pair": ms * 1e3 / B, "gbs": alg / ms / 1e6,
                      "frac_hbm": alg / ms / 1e6 / HBM, "tflops_useful": flops / ms / 1e9}), flush=True)


if __name__ == "__main__":
    for B in (1, 2, 4, 8, 16, 32, 64):
        point("conv4", B)
    for name in ("conv5", "conv3"):
        for B in (2, 16):
            point(name, B)
