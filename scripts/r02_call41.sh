#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 ) > gpurun_out/r02_c41_tests.log
cat gpurun_out/r02_c41_tests.log
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train > gpurun_out/r02_c41_bench.json 2> gpurun_out/r02_c41_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_c41_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["conv_ms_per_step"], d["e2e"]["value"], d["parity"]["ok"], d["parity"]["base_feat_max_rel_err"], d["gpu_launches"])
PY
