#!/bin/bash
# One gpurun call: the round's last experiments, each under its own timeout, results under gpurun_out/exp5_*.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/exp5_status.txt
timeout 240 python scripts/psroi_modes.py quick > gpurun_out/exp5_psroi_modes.jsonl 2> gpurun_out/exp5_psroi_modes.err
echo "psroi_modes exit $?" >> gpurun_out/exp5_status.txt
timeout 300 python bench.py --graph --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/exp5_bench_graph.json 2> gpurun_out/exp5_bench_graph.err
echo "bench --graph exit $?" >> gpurun_out/exp5_status.txt
cat gpurun_out/exp5_status.txt
