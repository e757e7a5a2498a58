#!/bin/bash
mkdir -p gpurun_out
( timeout 400 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py tests/test_train_conv_gpu.py tests/test_ops_gpu.py -q -x -k "not nms" 2>&1 | tail -6 ) > gpurun_out/r02_c34_tests.log
cat gpurun_out/r02_c34_tests.log
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train > gpurun_out/r02_c34_bench.json 2> gpurun_out/r02_c34_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_c34_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["conv_ms_per_step"], d["e2e"]["value"], d["parity"]["ok"], d["parity"]["base_feat_max_rel_err"], d["gpu_launches"])
PY
timeout 200 python scripts/engine_layer_times.py 16 > gpurun_out/r02_c34_layers.txt 2>&1
head -12 gpurun_out/r02_c34_layers.txt
