#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_train_engine_gpu.py tests/test_targets_gpu.py tests/test_train_gpu.py -q -x 2>&1 | tail -15 ) > gpurun_out/r02_c21_tests.log
cat gpurun_out/r02_c21_tests.log
timeout 600 python bench.py --train --steps 8 --warmup 4 > gpurun_out/r02_c21_train.json 2> gpurun_out/r02_c21_train.err
tail -3 gpurun_out/r02_c21_train.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_c21_train.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("ms_per_step", "engine_forward_ms", "engine_backward_ms", "heads_losses_optimizer_ms", "loss", "loss_finite", "launch")})
PY
