#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_ops_gpu.py -q -x -k "roi" 2>&1 | tail -8 ) > gpurun_out/r02_c63_tests.log
cat gpurun_out/r02_c63_tests.log
timeout 300 python bench.py --ops-only > gpurun_out/r02_c63_ops.json 2> gpurun_out/r02_c63_ops.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02_c63_ops.json").read().strip().splitlines()[-1])
    o = d.get("ops", d)
    for k in o:
        if k.startswith("roi_") or k.startswith("psroi_bwd"): print(k, json.dumps(o[k])[:300])
except Exception as e:
    print("failed", e); print(open("gpurun_out/r02_c63_ops.err").read()[-1500:])
PY
