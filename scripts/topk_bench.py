"""top-n_take of the RPN scores + gather: torch.sort (cub) + d2t_proposal_gather / one-CTA bitonic kernel / select + rank sort"""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'pytorch-detect-to-track_b200')
import torch
sys.argv = ['bench.py']
import bench
from d2t_b200 import ops
torch.cuda.set_device(0)
flush = torch.zeros(64 * 1024 * 1024, device='cuda')
for B, n_total, n_take in ((4, 28728, 6000), (4, 28728, 12000)):
    scores = torch.rand(B, n_total, device='cuda')
    boxes = torch.rand(B, n_total, 4, device='cuda') * 500
    def cub():
        order = torch.sort(scores, dim=1, descending=True, stable=True)[1]
        return ops.proposal_gather(boxes, scores, order, n_take)
    t = {"cub_sort+gather": bench.time_kernel(cub, 20, flush),
         "histogram select + rank sort (6 launches)": bench.time_kernel(lambda: ops.proposal_topk_gather(boxes, scores, n_take, split=True), 20, flush)}
    if n_take <= 8192:
        t["one_cta_bitonic"] = bench.time_kernel(lambda: ops.proposal_topk_gather(boxes, scores, n_take), 20, flush)
    print(B, n_total, n_take, {k: "%.1f us" % (v * 1e3) for k, v in t.items()}, flush=True)
