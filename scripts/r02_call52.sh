#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_detect_gpu.py -q -x -k "nms or detect" 2>&1 | tail -4 ) > gpurun_out/r02_c52_tests.log
cat gpurun_out/r02_c52_tests.log
timeout 300 python bench.py --ops-only > gpurun_out/r02_c52_ops.json 2> gpurun_out/r02_c52_ops.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_c52_ops.json").read().strip().splitlines()[-1])
d = d.get("ops", d)
for k, v in d.items():
    if k.startswith("nms") or k.startswith("detect"):
        print(k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items() if a != "note"})
PY
