#!/bin/bash
mkdir -p gpurun_out
( timeout 200 python -m pytest tests/test_ops_gpu.py tests/test_conv_gpu.py tests/test_model_gpu.py -q -x -k "corr or rfcn or engine_matches" 2>&1 | tail -4 ) > gpurun_out/r02_c53_tests.log
cat gpurun_out/r02_c53_tests.log
timeout 200 python bench.py --ops-only > gpurun_out/r02_c53_ops.json 2> gpurun_out/r02_c53_ops.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_c53_ops.json").read().strip().splitlines()[-1])
d = d.get("ops", d)
for k, v in d.items():
    if k.startswith("corr"):
        print(k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items() if a in ("ms", "ms_3xtf32", "ms_operator_api_nchw", "frac_hbm", "tflops_useful")})
PY
