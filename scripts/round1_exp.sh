#!/bin/bash
# One gpurun call: completion-counter hand-shake between consecutive conv launches (D2T_CONV_DONE=1), A/B on one box.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/exp6_status.txt
D2T_CONV_DONE=1 timeout 150 python -m pytest tests/test_conv_gpu.py -m gpu -x -q -k "engine" > gpurun_out/exp6_pytest_done.log 2>&1
echo "pytest engine (D2T_CONV_DONE=1) exit $?" >> gpurun_out/exp6_status.txt
D2T_CONV_DONE=1 timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/exp6_bench_done.json 2> gpurun_out/exp6_bench_done.err
echo "bench (D2T_CONV_DONE=1) exit $?" >> gpurun_out/exp6_status.txt
timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/exp6_bench_base.json 2> gpurun_out/exp6_bench_base.err
echo "bench (default) exit $?" >> gpurun_out/exp6_status.txt
cat gpurun_out/exp6_status.txt
tail -n 3 gpurun_out/exp6_pytest_done.log
python - <<'PY'
import json
for n in ("done", "base"):
    try:
        d = json.loads(open("gpurun_out/exp6_bench_%s.json" % n).read().strip().splitlines()[-1])
        print(n, d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["conv_ms_per_step"])
    except Exception as e:
        print(n, "no line:", e)
PY
