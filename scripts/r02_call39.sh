#!/bin/bash
mkdir -p gpurun_out
( timeout 120 python -m pytest tests/test_conv_gpu.py -q -x -k "stem" 2>&1 | tail -8 ) > gpurun_out/r02_c39_tests.log
cat gpurun_out/r02_c39_tests.log
( timeout 400 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py tests/test_train_engine_gpu.py -q -x 2>&1 | tail -4 ) > gpurun_out/r02_c39_tests2.log
cat gpurun_out/r02_c39_tests2.log
for m in 3 16; do
D2T_STEM_PASSES=$m timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train > gpurun_out/r02_c39_bench_stem$m.json 2> gpurun_out/r02_c39_bench_stem$m.err
done
python - <<'PY'
import json
for n in ("stem3", "stem16"):
    try:
        d = json.loads(open("gpurun_out/r02_c39_bench_%s.json" % n).read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d["roofline"]["conv_ms_per_step"], d["e2e"]["value"], d["parity"]["ok"], d["parity"]["base_feat_max_rel_err"], d["gpu_launches"])
    except Exception as e:
        print(n, "failed", e)
PY
