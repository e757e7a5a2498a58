#!/bin/bash
mkdir -p gpurun_out
( timeout 400 python -m pytest tests/test_train_conv_gpu.py tests/test_train_engine_gpu.py -q -x 2>&1 | tail -4 ) > gpurun_out/r02_c46_tests.log
cat gpurun_out/r02_c46_tests.log
for m in 0 1; do
D2T_CONV_ARES=$m timeout 300 python bench.py --train --steps 8 --warmup 4 > gpurun_out/r02_c46_train_ares$m.json 2> gpurun_out/r02_c46_train_ares$m.err
done
python - <<'PY'
import json
for m in (0, 1):
    d = json.loads(open("gpurun_out/r02_c46_train_ares%d.json" % m).read().strip().splitlines()[-1])
    print("ares", m, {k: round(d[k], 3) for k in ("ms_per_step", "engine_forward_ms", "engine_backward_ms", "heads_losses_optimizer_ms", "loss")})
PY
