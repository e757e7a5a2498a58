#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_ops_gpu.py -q -x -k "topk or proposal" 2>&1 | tail -6 ) > gpurun_out/r02_c45_tests.log
cat gpurun_out/r02_c45_tests.log
