"""Kernel timeline of ONE eval step (no CUDA graph) from torch.profiler: start / duration / stream of every kernel after the
trunk, to see what the tail of the step waits for."""
import os, sys, json, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-detect-to-track_b200")):
    sys.path.insert(0, p)
import torch
sys.argv = ["bench.py"]
import bench
from d2t_b200.engine import D2TEngine, GraphedEngine
from torch.profiler import profile, ProfilerActivity
torch.cuda.set_device(0)
H, W, pairs = bench.H, bench.W, 2
net = bench.build_net(101).cuda().eval()
im, info = bench.make_inputs(pairs, seed=1)
im, info = im.cuda(), info.cuda()
eng = D2TEngine(net, pairs, H, W)
if os.environ.get("TIMELINE_GRAPH", "1") == "1":
    eng = GraphedEngine(eng, pairs, H, W)
for _ in range(5):
    eng(im, info)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    eng(im, info)
    torch.cuda.synchronize()
f = tempfile.mktemp(suffix=".json")
prof.export_chrome_trace(f)
ev = [e for e in json.load(open(f))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
end = max(e["ts"] + e["dur"] for e in ev)
print("step: %d device ops, %.1f us from first start to last end" % (len(ev), end - t0))
last = -1.0
for i, e in enumerate(ev):
    name = e["name"]
    short = name[name.find("conv_igemm"):][:34] if "conv_igemm" in name else name.split("(")[0][-48:]
    if i >= len(ev) - int(sys.argv[1] if len(sys.argv) > 1 else 70):
        print("%8.1f us  +%6.1f  dur %7.1f  stream %3s  %s" % (e["ts"] - t0, e["ts"] - t0 - last if last >= 0 else 0.0, e["dur"], e["args"].get("stream"), short))
    last = max(last, e["ts"] + e["dur"] - t0)
