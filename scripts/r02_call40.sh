#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_train_engine_gpu.py -q -x -s -k "gradients_match" 2>&1 | grep -E "train engine vs|assert|Error|error" | head -20 ) > gpurun_out/r02_c40_tests.log
cat gpurun_out/r02_c40_tests.log
( D2T_STEM_PASSES=3 timeout 300 python -m pytest tests/test_train_engine_gpu.py -q -x -s -k "gradients_match" 2>&1 | grep -E "train engine vs|assert|Error|error|passed|failed" | head -20 ) > gpurun_out/r02_c40_tests3.log
cat gpurun_out/r02_c40_tests3.log
