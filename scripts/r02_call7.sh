#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_engine_gpu.py -q -s 2>&1 | tail -40 > gpurun_out/r02_c7_engine.log
timeout 900 python scripts/r02_bwd_profile.py > gpurun_out/r02_c7_bwd_profile.log 2>&1
tail -n 4 gpurun_out/r02_c7_engine.log; tail -n 40 gpurun_out/r02_c7_bwd_profile.log
