#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -q -x -m gpu 2>&1 | tail -5 ) > gpurun_out/r02_c61_tests.log
cat gpurun_out/r02_c61_tests.log
for m in 2 0; do
D2T_PSROI_BWD_INT=$m timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_c61_bench_$m.json 2> gpurun_out/r02_c61_bench_$m.err
done
python - <<'PY'
import json
for n in ("2", "0"):
    try:
        d = json.loads(open("gpurun_out/r02_c61_bench_%s.json" % n).read().strip().splitlines()[-1])
        print("bwd mode", n, d["value"], d["ms_per_step"], "train", d["train"]["ms_per_step"], d["train"].get("loss_finite"), "psroi_bwd", d["ops"]["psroi_bwd"]["ms"], d["ops"].get("psroi_bwd_fp64_tables", {}).get("ms"))
    except Exception as e:
        print(n, "failed", e)
PY
