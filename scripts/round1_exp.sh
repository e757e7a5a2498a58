#!/bin/bash
# One gpurun call: the round's last experiments, each under its own timeout, results under gpurun_out/exp3_*.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/exp3_status.txt
timeout 240 python scripts/psroi_modes.py > gpurun_out/exp3_psroi_modes.jsonl 2> gpurun_out/exp3_psroi_modes.err
echo "psroi_modes exit $?" >> gpurun_out/exp3_status.txt
timeout 200 python scripts/corr_sweep.py > gpurun_out/exp3_corr_sweep.jsonl 2> gpurun_out/exp3_corr_sweep.err
echo "corr_sweep exit $?" >> gpurun_out/exp3_status.txt
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/exp3_pytest.log 2>&1
echo "pytest (default = auto) exit $?" >> gpurun_out/exp3_status.txt
cat gpurun_out/exp3_status.txt
tail -n 4 gpurun_out/exp3_pytest.log
