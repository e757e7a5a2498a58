#!/bin/bash
mkdir -p gpurun_out
make -s -C pytorch-detect-to-track_b200/csrc trace > gpurun_out/r02_c31_make.log 2>&1
export D2T_B200_LIB=$PWD/pytorch-detect-to-track_b200/d2t_b200/libd2t_b200_trace.so
for cfg in "4 256 38 63 1024 1 1 0 1 16 res" "4 128 75 125 512 1 1 0 1 16 res"; do
  timeout 120 python scripts/conv_trace.py $cfg 2>&1 | grep -v Warn
done > gpurun_out/r02_c31_trace.log
cat gpurun_out/r02_c31_trace.log | cut -c1-220
