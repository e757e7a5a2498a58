#!/bin/bash
mkdir -p gpurun_out
( timeout 400 python -m pytest tests/test_train_engine_gpu.py -q -x 2>&1 | tail -4 ) > gpurun_out/r02_c48_tests.log
cat gpurun_out/r02_c48_tests.log
timeout 300 python bench.py --train --steps 8 --warmup 4 > gpurun_out/r02_c48_train.json 2> gpurun_out/r02_c48_train.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_c48_train.json").read().strip().splitlines()[-1])
print({k: round(d[k], 3) for k in ("ms_per_step", "engine_forward_ms", "engine_backward_ms", "heads_losses_optimizer_ms", "loss")}, d["gpu_launches"])
PY
