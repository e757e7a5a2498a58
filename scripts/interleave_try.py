"""Experiment: D2TEngine (2 pairs per launch, one chain) against D2TEngineStreams (2 chains of 1 pair on two streams,
enqueued layer by layer in turn, so the ~7.5 us dependent-launch gap of one chain hides behind the other's kernel).
Run once as is and once with D2T_CONV_PDL=0 (no programmatic dependent launch: the next kernel of a chain does not
sit on an SM waiting while the other chain could use it).  Run under `timeout`."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-detect-to-track_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from model.faster_rcnn.resnet import resnet
from d2t_b200.engine import D2TEngine, D2TEngineStreams, GraphedEngine

torch.manual_seed(3)
net = resnet(tuple(range(31)), 101, class_agnostic=True).create_architecture().cuda().eval()
H, W = 600, 1000
im = (torch.rand(2, 2, 3, H, W) * 256 - 128).cuda()
info = torch.tensor([H, W, 1.0]).view(1, 1, 3).expand(2, 2, 3).contiguous().cuda()
flush = torch.zeros(128 * 1024 * 1024, device="cuda")


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.add_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        tot += a.elapsed_time(b)
    return tot / n


res = {"pdl": os.environ.get("D2T_CONV_PDL", "1")}
e2 = D2TEngine(net, 2, H, W)
res["one_chain_ms"] = timeit(lambda: e2(im, info))
es = D2TEngineStreams(net, 2, H, W, chains=2)
res["two_chains_ms"] = timeit(lambda: es(im, info))
res["one_chain_again_ms"] = timeit(lambda: e2(im, info))
print(json.dumps(res), flush=True)
for name, eng in (("graph_one_chain", e2), ("graph_two_chains", es)):
    try:
        g = GraphedEngine(eng, 2, H, W)
        res[name + "_ms"] = timeit(lambda: g(im, info))
        o = g(im, info)
        torch.cuda.synchronize()
        r0 = eng(im, info)
        torch.cuda.synchronize()
        res[name + "_rois_equal"] = bool(torch.equal(o[0], r0[0]))
        res[name + "_cls_maxdiff"] = float((o[1] - r0[1]).abs().max())
        # replay determinism: 20 replays must reproduce the first bit for bit (a stale stream-K flag would not)
        first = [t.clone() for t in o[:4]]
        same = True
        for _ in range(20):
            o = g(im, info)
            torch.cuda.synchronize()
            same = same and all(torch.equal(a, b) for a, b in zip(first, o[:4]))
        res[name + "_replays_bit_identical"] = same
    except Exception as exc:
        res[name + "_error"] = repr(exc)[:300]
ref = e2(im, info)
out = es(im, info)
torch.cuda.synchronize()
res["rois_equal_frac"] = float((ref[0] == out[0]).all(-1).float().mean())
res["cls_prob_maxdiff"] = float((ref[1] - out[1]).abs().max())
res["bbox_maxdiff"] = float((ref[2] - out[2]).abs().max())
res["trk_maxdiff"] = float((ref[3] - out[3]).abs().max())
print(json.dumps(res))
