#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_conv_gpu.py tests/test_train_conv_gpu.py tests/test_train_engine_gpu.py tests/test_model_gpu.py -q 2>&1 | tail -15 > gpurun_out/r02_c13_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_c13_bench.json 2> gpurun_out/r02_c13_bench.err
make -s -C pytorch-detect-to-track_b200/csrc trace > gpurun_out/r02_c13_make.log 2>&1
export D2T_B200_LIB=$PWD/pytorch-detect-to-track_b200/d2t_b200/libd2t_b200_trace.so
for cfg in "4 256 38 63 1024 1 1 0 1 16 res" "4 64 150 250 256 1 1 0 1 16 res"; do
  timeout 120 python scripts/conv_trace.py $cfg 2>&1 | grep -v Warn
done > gpurun_out/r02_c13_trace.log
tail -n 5 gpurun_out/r02_c13_tests.log
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_c13_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["conv_ms_per_step"], d["e2e"]["value"], d["parity"])
print({k: d["train"][k] for k in ("ms_per_step", "engine_forward_ms", "engine_backward_ms", "heads_losses_optimizer_ms")})
PY
grep -E "^layer|^epilogue0|^mma" gpurun_out/r02_c13_trace.log | cut -c1-230
