#!/bin/bash
mkdir -p gpurun_out
( timeout 400 python -m pytest tests/test_train_conv_gpu.py tests/test_train_engine_gpu.py tests/test_conv_gpu.py -q -x -k "not chain_bit" 2>&1 | tail -4 ) > gpurun_out/r02_c49_tests.log
cat gpurun_out/r02_c49_tests.log
for m in 0 1; do
D2T_WGRAD_REDUCE=$m timeout 300 python bench.py --train --steps 8 --warmup 4 > gpurun_out/r02_c49_train_$m.json 2> gpurun_out/r02_c49_train_$m.err
done
python - <<'PY'
import json
for m in (0, 1):
    d = json.loads(open("gpurun_out/r02_c49_train_%d.json" % m).read().strip().splitlines()[-1])
    print("reduce", m, {k: round(d[k], 3) for k in ("ms_per_step", "engine_forward_ms", "engine_backward_ms", "heads_losses_optimizer_ms", "loss")}, d["gpu_launches"])
PY
