#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py tests/test_detect_gpu.py -q -x 2>&1 | tail -5 ) > gpurun_out/r02_c22_tests.log
cat gpurun_out/r02_c22_tests.log
D2T_ENGINE_FORK=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train > gpurun_out/r02_c22_bench_nofork.json 2> gpurun_out/r02_c22_bench_nofork.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train > gpurun_out/r02_c22_bench_fork.json 2> gpurun_out/r02_c22_bench_fork.err
tail -2 gpurun_out/r02_c22_bench_fork.err
python - <<'PY'
import json
for n in ("nofork", "fork"):
    try:
        d = json.loads(open("gpurun_out/r02_c22_bench_%s.json" % n).read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d["roofline"]["conv_ms_per_step"], d["e2e"]["value"], d["parity"]["ok"], d["parity"]["base_feat_max_rel_err"], d["gpu_launches"])
    except Exception as e:
        print(n, "failed", e)
PY
