#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py tests/test_detect_gpu.py -q -x -k "engine or rfcn or detect" 2>&1 | tail -4 ) > gpurun_out/r02_c42_tests.log
cat gpurun_out/r02_c42_tests.log
for m in 2 1; do
D2T_ENGINE_FORK=$m timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train > gpurun_out/r02_c42_bench_fork$m.json 2> gpurun_out/r02_c42_bench_fork$m.err
done
python - <<'PY'
import json
for n in ("fork2", "fork1"):
    try:
        d = json.loads(open("gpurun_out/r02_c42_bench_%s.json" % n).read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d["roofline"]["conv_ms_per_step"], d["e2e"]["value"], d["parity"]["ok"], d["parity"]["identical_proposals_frac"], d["gpu_launches"])
    except Exception as e:
        print(n, "failed", e)
PY
