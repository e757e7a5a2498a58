#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_engine_gpu.py tests/test_train_gpu.py -q -s 2>&1 | tail -40 > gpurun_out/r02_c9_engine.log
timeout 900 python bench.py --train --steps 10 --warmup 4 > gpurun_out/r02_c9_train.json 2> gpurun_out/r02_c9_train.err
echo "train rc=$?" >> gpurun_out/r02_c9_train.err
# A/B: conv chain with / without the mask code in the epilogue, same box
timeout 600 python bench.py --steps 10 --warmup 3 --no-train --no-cpu-baseline > gpurun_out/r02_c9_bench_a.json 2> gpurun_out/r02_c9_bench_a.err
D2T_B200_LIB=$PWD/pytorch-detect-to-track_b200/d2t_b200/libd2t_b200_ab.so timeout 600 python bench.py --steps 10 --warmup 3 --no-train --no-cpu-baseline > gpurun_out/r02_c9_bench_b.json 2> gpurun_out/r02_c9_bench_b.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-train --no-cpu-baseline > gpurun_out/r02_c9_bench_a2.json 2> gpurun_out/r02_c9_bench_a2.err
grep -E "passed|failed|median|losses" gpurun_out/r02_c9_engine.log; tail -n 2 gpurun_out/r02_c9_train.err
python - <<'PY'
import json
for f in ("a", "b", "a2"):
    try:
        d = json.loads(open("gpurun_out/r02_c9_bench_%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["roofline"]["conv_ms_per_step"])
    except Exception as e:
        print(f, "failed", e)
d = json.loads(open("gpurun_out/r02_c9_train.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("ms_per_step", "engine_forward_ms", "engine_backward_ms", "heads_losses_optimizer_ms", "loss", "launch")})
PY
