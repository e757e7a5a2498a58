#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py -q -x 2>&1 | tail -5 ) > gpurun_out/r02_c20_tests.log
cat gpurun_out/r02_c20_tests.log
D2T_CONV_EARLY_B=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train > gpurun_out/r02_c20_bench_noearly.json 2> gpurun_out/r02_c20_bench_noearly.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train > gpurun_out/r02_c20_bench_early.json 2> gpurun_out/r02_c20_bench_early.err
python - <<'PY'
import json
for n in ("noearly", "early"):
    try:
        d = json.loads(open("gpurun_out/r02_c20_bench_%s.json" % n).read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d["roofline"]["conv_ms_per_step"], d["e2e"]["value"], d["parity"]["ok"], d["parity"]["base_feat_max_rel_err"], d["gpu_launches"])
    except Exception as e:
        print(n, "failed", e)
PY
