#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_engine_gpu.py tests/test_train_conv_gpu.py tests/test_conv_gpu.py -q 2>&1 | tail -15 > gpurun_out/r02_c10_tests.log
timeout 900 python scripts/r02_heads_profile.py > gpurun_out/r02_c10_heads_profile.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_c10_bench.json 2> gpurun_out/r02_c10_bench.err
tail -n 6 gpurun_out/r02_c10_tests.log
