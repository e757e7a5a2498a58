#!/bin/bash
mkdir -p gpurun_out
( timeout 200 python -m pytest tests/test_ops_gpu.py -q -x -k "topk or proposal" 2>&1 | tail -8 ) > gpurun_out/r02_c44_tests.log
cat gpurun_out/r02_c44_tests.log
timeout 120 python - <<'PY' 2>&1 | grep -v Warn
import sys, torch
sys.path.insert(0, "pytorch-detect-to-track_b200")
from d2t_b200 import ops
for (n_total, n_take, B) in [(28728, 6000, 4), (28728, 12000, 2)]:
    scores = torch.rand(B, n_total, device="cuda"); boxes = torch.rand(B, n_total, 4, device="cuda")
    def a():
        order = torch.sort(scores, dim=1, descending=True, stable=True)[1]
        return ops.proposal_gather(boxes, scores, order, n_take)
    def b():
        return ops.proposal_topk_gather(boxes, scores, n_take)
    for name, fn in (("torch.sort + gather", a), ("topk_gather", b)):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        print(n_total, n_take, B, name, "%.1f us" % (e0.elapsed_time(e1) / 20 * 1e3))
PY
