#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r02_c14_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train > gpurun_out/r02_c14_bench.json 2> gpurun_out/r02_c14_bench.err
tail -n 8 gpurun_out/r02_c14_tests.log
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_c14_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["conv_ms_per_step"], d["e2e"]["value"], d["parity"]["ok"], d["parity"]["base_feat_max_rel_err"])
for k in ("corr_conv4", "corr_conv5", "corr_conv3"):
    print(k, {a: d["ops"][k][a] for a in ("ms", "ms_3xtf32", "frac_hbm", "tflops_useful", "ms_operator_api_nchw")})
PY
