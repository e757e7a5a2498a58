#!/bin/bash
mkdir -p gpurun_out
( timeout 200 python -m pytest tests/test_conv_gpu.py -q -x -k "a_resident or conv_matches_torch or engine_matches" 2>&1 | tail -6 ) > gpurun_out/r02_c32_tests.log
cat gpurun_out/r02_c32_tests.log
for m in 0 1; do
D2T_CONV_ARES=$m timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train > gpurun_out/r02_c32_bench_ares$m.json 2> gpurun_out/r02_c32_bench_ares$m.err
done
python - <<'PY'
import json
for n in ("ares0", "ares1"):
    try:
        d = json.loads(open("gpurun_out/r02_c32_bench_%s.json" % n).read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d["roofline"]["conv_ms_per_step"], d["e2e"]["value"], d["parity"]["ok"], d["parity"]["base_feat_max_rel_err"], d["gpu_launches"])
    except Exception as e:
        print(n, "failed", e)
PY
timeout 200 python scripts/engine_layer_times.py 16 > gpurun_out/r02_c32_layers_ares1.txt 2>&1
grep "res\|256 -> 512\|512 -> 1024" gpurun_out/r02_c32_layers_ares1.txt | head -8
