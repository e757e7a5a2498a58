#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py -q -x -k "correlation or rfcn" 2>&1 | tail -6 ) > gpurun_out/r02_c36_tests.log
cat gpurun_out/r02_c36_tests.log
timeout 300 python - <<'PY' 2>&1 | grep -v Warn
import sys, os, torch
sys.path.insert(0, "pytorch-detect-to-track_b200")
from d2t_b200 import ops
a, b = torch.randn(2, 1024, 38, 63, device="cuda"), torch.randn(2, 1024, 38, 63, device="cuda")
g = torch.randn(2, 289, 38, 63, device="cuda")
for tc in (True, False):
    ops.TENSOR_CORE_CORRELATION = tc
    for _ in range(3): ops.correlation_backward(a, b, g, 8, 1, 8, 1, 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): ops.correlation_backward(a, b, g, 8, 1, 8, 1, 1)
    e1.record(); torch.cuda.synchronize()
    print("correlation_backward conv4 B=2, tensor cores" if tc else "correlation_backward conv4 B=2, SIMT gather", e0.elapsed_time(e1) / 5, "ms")
PY
