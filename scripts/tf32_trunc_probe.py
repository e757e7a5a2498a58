"""Does tcgen05 kind::tf32 ignore the low 13 mantissa bits of its fp32 operands (truncate)?
Run the single-pass conv on x and on trunc13(x) (weights likewise): bit-identical outputs mean the tensor core
reads an fp32 operand as its truncation, which is what the 3xTF32 split (lo = x - trunc13(x)) relies on."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-detect-to-track_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from d2t_b200 import conv as dc
torch.manual_seed(0)
N, Cin, H, W, Cout = 2, 256, 38, 63, 256
x = torch.randn(N, Cin, H, W, device="cuda")
w = torch.randn(Cout, Cin, 3, 3, device="cuda") * 0.02


def trunc13(t):
    return (t.view(torch.int32) & ~0x1fff).view(torch.float32)


outs = []
for mode in ("full", "trunc"):
    xs = dc.ActTensor.from_nchw(trunc13(x) if mode == "trunc" else x)
    layer = dc.ConvLayer(xs, trunc13(w) if mode == "trunc" else w, None, None, 1, 1, 1, False, None, passes=1, want_nchw=True)
    layer.run()
    torch.cuda.synchronize()
    outs.append(layer.out_nchw.clone())
print("bit-identical:", bool(torch.equal(outs[0], outs[1])), " max diff %.3e" % float((outs[0] - outs[1]).abs().max()))
