#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_ops_gpu.py -q -x -k "nms or proposal" 2>&1 | tail -4 )
( timeout 600 python -m pytest tests/test_model_gpu.py tests/test_detect_gpu.py -q -x 2>&1 | tail -3 )
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r02_c68_bench.json 2> gpurun_out/r02_c68_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_c68_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["parity"]["ok"], d["parity"]["identical_proposals_frac"], "train", d["train"]["ms_per_step"])
for k in ("nms_6000_x4img", "nms_12000_x4img"): print(k, d["ops"][k])
PY
