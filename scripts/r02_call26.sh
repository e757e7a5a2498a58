#!/bin/bash
mkdir -p gpurun_out
( timeout 200 python -m pytest tests/test_conv_gpu.py -q -x -k "pair" 2>&1 | tail -4 ) > gpurun_out/r02_c26_tests.log
cat gpurun_out/r02_c26_tests.log
D2T_CONV_PAIR=2 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train > gpurun_out/r02_c26_bench_pair2.json 2> gpurun_out/r02_c26_bench_pair2.err
python - <<'PY'
import json
for n in ("pair2",):
    try:
        d = json.loads(open("gpurun_out/r02_c26_bench_%s.json" % n).read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d["roofline"]["conv_ms_per_step"], d["e2e"]["value"], d["parity"]["ok"], d["parity"]["base_feat_max_rel_err"], d["gpu_launches"])
    except Exception as e:
        print(n, "failed", e)
PY
