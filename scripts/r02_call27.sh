#!/bin/bash
mkdir -p gpurun_out
( timeout 120 python -m pytest tests/test_conv_gpu.py -q -x -k "a_resident" 2>&1 | tail -12 ) > gpurun_out/r02_c27_tests.log
cat gpurun_out/r02_c27_tests.log
