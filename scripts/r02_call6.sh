#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_engine_gpu.py -q -s 2>&1 | tail -60 > gpurun_out/r02_c6_engine.log
timeout 900 python -m pytest tests/test_model_gpu.py -q -s 2>&1 | tail -40 > gpurun_out/r02_c6_model.log
timeout 600 python scripts/r02_train_debug.py > gpurun_out/r02_c6_debug.log 2>&1
tail -n 4 gpurun_out/r02_c6_engine.log; tail -n 4 gpurun_out/r02_c6_model.log
