#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py -q -x 2>&1 | tail -4 )
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-train > gpurun_out/r02_c65_bench.json 2> gpurun_out/r02_c65_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_c65_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["conv_ms_per_step"], d["e2e"]["value"], d["parity"])
for k in ("psroi_vote_cls_softmax", "psroi_vote_bbox"): print(k, d["ops"][k])
PY
