#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py -q -x -k "chain" 2>&1 | tail -25 > gpurun_out/r02_c15_tests.log
cat gpurun_out/r02_c15_tests.log
D2T_CONV_CHAIN=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train > gpurun_out/r02_c15_bench_nochain.json 2> gpurun_out/r02_c15_bench_nochain.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train > gpurun_out/r02_c15_bench_chain.json 2> gpurun_out/r02_c15_bench_chain.err
tail -3 gpurun_out/r02_c15_bench_chain.err
python - <<'PY'
import json
for n in ("nochain", "chain"):
    try:
        d = json.loads(open("gpurun_out/r02_c15_bench_%s.json" % n).read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d["roofline"]["conv_ms_per_step"], d["e2e"]["value"], d["parity"]["ok"], d["parity"]["base_feat_max_rel_err"], d["gpu_launches"])
    except Exception as e:
        print(n, "failed", e)
PY
