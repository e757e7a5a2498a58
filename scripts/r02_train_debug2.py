"""pinpoint the first diverging gradient tensor of the training engine"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-detect-to-track_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import torch.nn.functional as F
import test_train_engine_gpu as T
from d2t_b200.train import D2TTrainEngine

B, H, W = 1, 160, 224
net, im_data, im_info, gt, nb = T._setup(50, B, H, W)
eng = D2TTrainEngine(net, B, H, W)
out, loss = eng.forward_backward(im_data, im_info, gt, nb)
torch.cuda.synchronize()
N = 2 * B
frames = im_data.permute(1, 0, 2, 3, 4).reshape(N, *im_data.shape[2:]).contiguous()
b = net.RFCN_base
conv1 = b[3](b[2](b[1](b[0](frames))))
conv2 = b[4](conv1)
conv3 = b[5](conv2)
conv4 = b[6](conv3)
conv5 = b[7](conv4)
pre_base = b[8](conv5)
base = F.relu(pre_base)
cls_map, bbox_map = net.RFCN_cls_net(base), net.RFCN_bbox_net(base)
rpn = net.RFCN_rpn
pre_rc = rpn.RPN_Conv(base)
rc = F.relu(pre_rc)
score, delta = rpn.RPN_cls_score(rc), rpn.RPN_bbox_pred(rc)
trk = net._tracking_maps(conv3, conv4, conv5, bbox_map, B)
for t in (pre_rc, pre_base, conv5, conv4, conv3, rc, base, bbox_map):
    t.retain_grad()
torch.autograd.backward([cls_map, bbox_map, score, delta, trk], [g.clone() for g in eng.leaf_grads])
torch.cuda.synchronize()
rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
print("forward: rc %.2e base %.2e conv5 %.2e" % (rel(eng._meta[id(rpn.RPN_Conv.weight)][0].out.to_nchw(), rc), rel(eng.base_feat.to_nchw(), base), rel(eng.feat_nchw[7], conv5)))
g_rc = eng.g_rc.to_nchw()
print("g_rc    vs d/d(pre_rc)   %.3e   (scale %.3e, mine max %.3e)" % (rel(g_rc, pre_rc.grad), float(pre_rc.grad.abs().max()), float(g_rc.abs().max())))
# unmasked comparison: where do they differ?
d = (g_rc - pre_rc.grad).abs()
print("   mismatching elements > 1e-3 of scale: %d of %d; of those with rc == 0: %d" % (int((d > 1e-3 * pre_rc.grad.abs().max()).sum()), d.numel(), int(((d > 1e-3 * pre_rc.grad.abs().max()) & (rc == 0)).sum())))
g_base = eng.g_base.to_nchw()
print("g_base  vs d/d(pre_base) %.3e   (scale %.3e)" % (rel(g_base, pre_base.grad), float(pre_base.grad.abs().max())))
g5 = eng.g_conv5.to_nchw()
# engine g_conv5 = gradient w.r.t. the pre-activation of layer4's last block = conv5.grad masked by conv5 > 0
want5 = conv5.grad * (conv5 > 0)
print("g_conv5 vs masked d/d(conv5) %.3e (scale %.3e)" % (rel(g5, want5), float(want5.abs().max())))
print("g_bbox  vs d/d(bbox_map) %.3e" % rel(eng.g_bbox.to_nchw(), bbox_map.grad))
for tag, ref in ((7, conv5), (6, conv4), (5, conv3)):
    pass
print("amax: g_score %.3e g_delta %.3e g_rc %.3e (true %.3e) g_base %.3e (true %.3e)" % (float(eng.g_score.amax), float(eng.g_delta.amax), float(eng.g_rc.amax), float(g_rc.abs().max()), float(eng.g_base.amax), float(g_base.abs().max())))
# the same dgrad chain re-run stand-alone from torch's tensors
from d2t_b200 import conv as dc
gs = dc.ActTensor.from_nchw(eng.leaf_grads[2].contiguous())
gd = dc.ActTensor.from_nchw(eng.leaf_grads[3].contiguous())
rcs = dc.ActTensor.from_nchw(rc.detach().contiguous(), cstride=512)
o = dc.ActTensor(N, rc.size(2), rc.size(3), 512, cstride=512)
am = lambda w: w.detach().abs().max().reshape(1).float()
a = dc.DgradConv(gs, rpn.RPN_cls_score.weight, None, 0, 1, am(rpn.RPN_cls_score.weight), out=o, mask=rcs)
bq = dc.DgradConv(gd, rpn.RPN_bbox_pred.weight, None, 0, 1, am(rpn.RPN_bbox_pred.weight), out=o, residual=o, mask=rcs)
a.run(); bq.run()
torch.cuda.synchronize()
print("stand-alone chain vs d/d(pre_rc) %.3e ; vs engine g_rc %.3e" % (rel(o.to_nchw(), pre_rc.grad), rel(o.to_nchw(), g_rc)))
