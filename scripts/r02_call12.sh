#!/bin/bash
mkdir -p gpurun_out
make -s -C pytorch-detect-to-track_b200/csrc trace > gpurun_out/r02_c12_make.log 2>&1
export D2T_B200_LIB=$PWD/pytorch-detect-to-track_b200/d2t_b200/libd2t_b200_trace.so
for cfg in "4 256 38 63 1024 1 1 0 1 16 res" "4 1024 38 63 256 1 1 0 1 16" "4 256 38 63 256 3 1 1 1 16" "4 512 38 63 512 3 1 2 2 16" "4 64 150 250 256 1 1 0 1 16 res" "4 128 75 125 512 1 1 0 1 16 res"; do
  timeout 120 python scripts/conv_trace.py $cfg 2>&1 | grep -v Warn
  echo
done > gpurun_out/r02_c12_trace.log
tail -n 50 gpurun_out/r02_c12_trace.log
