"""per-step device times of the training engine's backward pass (Res-101, 600x1000, 2 pairs), grouped by kind"""
import os, sys, json, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-detect-to-track_b200")):
    sys.path.insert(0, p)
import torch
sys.argv = ["bench.py"]
import bench
from d2t_b200 import synth, conv as dc
from d2t_b200.train import D2TTrainEngine
torch.cuda.set_device(0)
H, W, pairs = bench.H, bench.W, 2
net = bench.build_net(101).cuda()
im, info = bench.make_inputs(pairs, seed=1)
im, info = im.cuda(), info.cuda()
synth.calibrate_batchnorm(net, bench.make_inputs(1, seed=1)[0].view(2, 3, H, W).cuda())
net.train()
gt = torch.from_numpy(synth.make_gt_boxes(pairs, 30, seed=2, height=H, width=W)).cuda()
nb = (gt[..., 4] > 0).sum(-1, keepdim=True)
eng = D2TTrainEngine(net, pairs, H, W)
for _ in range(3):
    out, loss = eng.forward_backward(im, info, gt, nb)
torch.cuda.synchronize()
# whole phases
ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
from model.rpn.proposal_target_layer_cascade import train_heads
ev[0].record()
with torch.no_grad():
    i2 = eng._begin(im, info)
    for layer in eng.layers + eng.corr_layers + [eng.trk_layer]:
        layer.run()
ev[1].record()
leaves = [t.detach().requires_grad_() for t in (eng.cls_map, eng.bbox_map, eng.rpn_score, eng.rpn_delta, eng.trk_layer.out_nchw)]
o = train_heads(net, pairs, None, None, None, None, leaves[0], leaves[1], i2, gt, nb, rpn_maps=(leaves[2], leaves[3]), trk_map=leaves[4])
l = o[4].mean() + o[5].mean() + o[6].mean() + o[7].mean() + o[9].mean()
ev[2].record()
grads = torch.autograd.grad(l, leaves, allow_unused=True)
ev[3].record()
eng._run_backward()
ev[4].record()
torch.cuda.synchronize()
print(json.dumps({"forward_ms": ev[0].elapsed_time(ev[1]), "heads_forward_ms": ev[1].elapsed_time(ev[2]),
                  "heads_autograd_ms": ev[2].elapsed_time(ev[3]), "engine_backward_ms": ev[3].elapsed_time(ev[4])}))
import time
t0 = time.time(); torch.cuda.synchronize()
o = train_heads(net, pairs, None, None, None, None, leaves[0], leaves[1], i2, gt, nb, rpn_maps=(leaves[2], leaves[3]), trk_map=leaves[4])
torch.cuda.synchronize(); print("heads forward wall ms", (time.time() - t0) * 1e3)
# per step
groups = collections.OrderedDict()
rows = []
for fn, off, label in eng.bwd:
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record(); fn(); b.record(); b.synchronize()
    ms = a.elapsed_time(b)
    rows.append((label, ms))
    kind = label.split(" ")[0]
    groups[kind] = groups.get(kind, 0.0) + ms
print(json.dumps(groups))
agg = collections.OrderedDict()
for label, ms in rows:
    c, t = agg.get(label, (0, 0.0))
    agg[label] = (c + 1, t + ms)
for label, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-44s x%-3d total %8.3f ms  each %7.3f ms" % (label, c, t, t / c))
# inside one wgrad: pack x / pack g / gemm
wl = max(eng.wgrads, key=lambda w: w.flops)
for wl in (eng.wgrads[len(eng.wgrads) // 2], eng.wgrads[5], eng.wgrads[-1]):
    x, g, sc = wl.x, wl.g, wl.scratch
    O, I = wl.grad_w.shape[:2]
    st = torch.cuda.current_stream().cuda_stream
    from d2t_b200._lib import lib
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    torch.cuda.synchronize()
    evs[0].record()
    lib().d2t_wgrad_pack_input(x.x.data_ptr(), x.N, x.H, x.W, x.cstride, I, wl.stride, wl.xh, wl.xw, wl.xp, sc.xt.data_ptr(), st)
    evs[1].record()
    lib().d2t_wgrad_pack_grad(g.x.data_ptr(), g.N, g.H, g.W, g.cstride, O, wl.gp, wl.S, wl.dil, wl.pad, g.amax.data_ptr(), sc.g_hi.data_ptr(), sc.g_lo.data_ptr(), st)
    evs[2].record()
    lib().d2t_conv_plan_run(wl.plan, st)
    evs[3].record()
    torch.cuda.synchronize()
    print("wgrad", tuple(wl.grad_w.shape), "@%dx%d" % (g.H, g.W), "pack_x %.3f pack_g %.3f gemm %.3f ms; %.1f TF/s useful" % (
        evs[0].elapsed_time(evs[1]), evs[1].elapsed_time(evs[2]), evs[2].elapsed_time(evs[3]), wl.flops / evs[2].elapsed_time(evs[3]) / 1e9))
