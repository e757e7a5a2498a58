#!/bin/bash
mkdir -p gpurun_out
( timeout 150 python -m pytest tests/test_conv_gpu.py -q -x -k "pair" 2>&1 | tail -15 ) > gpurun_out/r02_c24_tests.log
cat gpurun_out/r02_c24_tests.log
