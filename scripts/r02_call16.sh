#!/bin/bash
mkdir -p gpurun_out
( timeout 120 python -m pytest tests/test_conv_gpu.py -q -x -k "chain" 2>&1 | tail -15 ) > gpurun_out/r02_c16_tests.log
cat gpurun_out/r02_c16_tests.log
( timeout 100 python scripts/chain_bench.py 8
  D2T_CHAIN_NOSYNC=1 timeout 100 python scripts/chain_bench.py 8
  D2T_CHAIN_COOP=0 timeout 100 python scripts/chain_bench.py 8 ) 2>&1 | grep -v Warn | tee gpurun_out/r02_c16_chain_bench.txt
