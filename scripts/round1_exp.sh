#!/bin/bash
# One gpurun call: the round's last experiments, each under its own timeout, results under gpurun_out/exp4_*.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/exp4_status.txt
timeout 240 python scripts/psroi_modes.py quick > gpurun_out/exp4_psroi_modes.jsonl 2> gpurun_out/exp4_psroi_modes.err
echo "psroi_modes exit $?" >> gpurun_out/exp4_status.txt
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/exp4_pytest.log 2>&1
echo "pytest (default = auto) exit $?" >> gpurun_out/exp4_status.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/exp4_bench.json 2> gpurun_out/exp4_bench.err
echo "bench exit $?" >> gpurun_out/exp4_status.txt
cat gpurun_out/exp4_status.txt
tail -n 4 gpurun_out/exp4_pytest.log
