"""diagnostics for the training engine: per-parameter gradient error table; packed-weight freshness after an update"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-detect-to-track_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import test_train_engine_gpu as T
from d2t_b200.train import D2TTrainEngine

B, H, W = 1, 160, 224
net, im_data, im_info, gt, nb = T._setup(50, B, H, W)
eng = D2TTrainEngine(net, B, H, W)
out, loss = eng.forward_backward(im_data, im_info, gt, nb)
torch.cuda.synchronize()
mine = eng.flat.clone()
eng.flat.zero_()
T._torch_param_grads(net, im_data, B, [g.clone() for g in eng.leaf_grads])
torch.cuda.synchronize()
ref = eng.flat.clone()
names = {id(p): n for n, p in net.named_parameters()}
for p in eng.params:
    o = p.grad.storage_offset()
    a, b = mine[o:o + p.numel()], ref[o:o + p.numel()]
    sc = float(b.abs().max())
    print("%-50s scale %.3e  err %.2e  mine_max %.3e" % (names[id(p)], sc, float((a - b).abs().max()) / max(sc, 1e-30), float(a.abs().max())))
print("leaf grad maxima", [float(g.abs().max()) for g in eng.leaf_grads])
# ---- freshness of the packed operands
opt = torch.optim.SGD(eng.params, lr=1e-6, momentum=0.9)
eng.flat.copy_(mine)
opt.step()
eng.refresh_weights()
torch.cuda.synchronize()
import math
bad = 0
for l in [eng.trk_layer] + eng.layers:
    w = l.weight.detach()
    O, I, R, S = w.shape
    amax = float(l.amax_w)
    true = float(w.abs().max())
    eb = max(15, min(254, (int(torch.tensor(amax).view(torch.int32)) >> 23) & 0xff)) if False else None
    e = 141 - max(15, min(254, (torch.tensor([amax]).view(torch.int32).item() >> 23) & 0xff))
    rec = (l.w_hi.float() + l.w_lo.float()) * 2.0 ** (-e)
    Ip = rec.shape[1] // (R * S)
    rec = rec.view(O, R * S, Ip)[:, :, :I].permute(0, 2, 1).reshape(O, I, R, S)
    err = float((rec - w).abs().max()) / true
    if err > 1e-6 or abs(amax - true) > 0:
        bad += 1
        print("STALE?", tuple(w.shape), "amax", amax, "true", true, "packed err", err)
print("layers with stale operands:", bad)
net.eval()
with torch.no_grad():
    frames = im_data.permute(1, 0, 2, 3, 4).reshape(2 * B, 3, H, W).contiguous()
    c3, c4, c5, base = net._im_to_head(frames)
    eng._begin(im_data, im_info)
    for layer in eng.layers:
        layer.run()
rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
print("after update: conv3 %.2e conv4 %.2e conv5 %.2e base %.2e" % (rel(eng.feat_nchw[5], c3), rel(eng.feat_nchw[6], c4), rel(eng.feat_nchw[7], c5), rel(eng.base_feat.to_nchw(), base)))
