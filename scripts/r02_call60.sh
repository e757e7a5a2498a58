#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_ops_gpu.py -q -x -k "psroi" 2>&1 | tail -4 )
for d in 0 1; do echo diag $d; D2T_BWD_DIAG=$d timeout 200 python scripts/psroi_bwd_bench.py 2>&1 | grep '"limb"' | cut -c1-110; done
