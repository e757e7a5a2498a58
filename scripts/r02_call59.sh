#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_ops_gpu.py -q -x -k "psroi" 2>&1 | tail -8 ) > gpurun_out/r02_c59_tests.log
cat gpurun_out/r02_c59_tests.log
timeout 200 python scripts/psroi_bwd_bench.py > gpurun_out/r02_c59_bwd.jsonl 2> gpurun_out/r02_c59_bwd.err
cat gpurun_out/r02_c59_bwd.jsonl | cut -c1-150; tail -3 gpurun_out/r02_c59_bwd.err
