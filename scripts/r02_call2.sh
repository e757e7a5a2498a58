#!/bin/bash
# round 2, GPU call 2: backward-data / weight-gradient unit tests, the training engine, the training bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_conv_gpu.py -q -s 2>&1 | tail -150 > gpurun_out/r02_c2_conv.log
timeout 900 python -m pytest tests/test_train_engine_gpu.py -q -s 2>&1 | tail -80 > gpurun_out/r02_c2_engine.log
timeout 900 python bench.py --train --steps 5 --warmup 3 > gpurun_out/r02_c2_train.json 2> gpurun_out/r02_c2_train.err
echo "train rc=$?" >> gpurun_out/r02_c2_train.err
tail -n 3 gpurun_out/r02_c2_conv.log; tail -n 3 gpurun_out/r02_c2_engine.log; tail -n 3 gpurun_out/r02_c2_train.err
