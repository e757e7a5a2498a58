#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_ops_gpu.py -q -x -k "psroi" 2>&1 | tail -4 )
timeout 200 python scripts/psroi_bwd_bench.py 2>&1 | grep -v int_mc | cut -c1-150
