"""torch.profiler over the eager part of the training step (heads: proposal step, target layers, PSRoI, losses, autograd)"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-detect-to-track_b200")):
    sys.path.insert(0, p)
import torch
sys.argv = ["bench.py"]
import bench
from d2t_b200 import synth
from d2t_b200.train import D2TTrainEngine
from torch.profiler import profile, ProfilerActivity
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
opt = torch.optim.SGD(eng.params, lr=1e-5, momentum=0.9, weight_decay=1e-4, fused=True)
for _ in range(4):
    out, loss = eng.forward_backward(im, info, gt, nb)
    opt.step(); eng.refresh_weights()
torch.cuda.synchronize()
import time
t0 = time.time()
for _ in range(5):
    out, loss = eng.forward_backward(im, info, gt, nb)
    opt.step(); eng.refresh_weights()
torch.cuda.synchronize()
print("wall ms/step", (time.time() - t0) / 5 * 1e3)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        out, loss = eng.forward_backward(im, info, gt, nb)
        opt.step(); eng.refresh_weights()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=60))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=60))
