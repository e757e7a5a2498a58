#!/bin/bash
mkdir -p gpurun_out
( timeout 200 python -m pytest tests/test_conv_gpu.py -q -x -k "pair" 2>&1 | tail -8 ) > gpurun_out/r02_c25_tests.log
cat gpurun_out/r02_c25_tests.log
for m in 0 1 2; do
D2T_CONV_PAIR=$m timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train > gpurun_out/r02_c25_bench_pair$m.json 2> gpurun_out/r02_c25_bench_pair$m.err
done
python - <<'PY'
import json
for n in ("pair0", "pair1", "pair2"):
    try:
        d = json.loads(open("gpurun_out/r02_c25_bench_%s.json" % n).read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d["roofline"]["conv_ms_per_step"], d["e2e"]["value"], d["parity"]["ok"], d["parity"]["base_feat_max_rel_err"], d["gpu_launches"])
    except Exception as e:
        print(n, "failed", e)
PY
D2T_CONV_PAIR=1 timeout 200 python scripts/engine_layer_times.py 16 > gpurun_out/r02_c25_layers_pair1.txt 2>&1
timeout 200 python scripts/engine_layer_times.py 16 > gpurun_out/r02_c25_layers_pair0.txt 2>&1
