#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_detect_gpu.py tests/test_model_gpu.py -q -x -k "nms or detect or proposal or engine or rfcn" 2>&1 | tail -5 ) > gpurun_out/r02_c23_tests.log
cat gpurun_out/r02_c23_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train > gpurun_out/r02_c23_bench.json 2> gpurun_out/r02_c23_bench.err
tail -2 gpurun_out/r02_c23_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_c23_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["conv_ms_per_step"], d["e2e"]["value"], d["parity"]["ok"], d["gpu_launches"])
print({k: v for k, v in d["ops"].items() if k.startswith("nms")})
PY
timeout 300 python bench.py --train --steps 8 --warmup 4 > gpurun_out/r02_c23_train.json 2> gpurun_out/r02_c23_train.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_c23_train.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("ms_per_step", "engine_forward_ms", "engine_backward_ms", "heads_losses_optimizer_ms", "loss", "loss_finite")})
PY
