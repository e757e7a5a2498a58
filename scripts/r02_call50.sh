#!/bin/bash
mkdir -p gpurun_out
make -s -C pytorch-detect-to-track_b200/csrc trace > gpurun_out/r02_c50_make.log 2>&1
export D2T_B200_LIB=$PWD/pytorch-detect-to-track_b200/d2t_b200/libd2t_b200_trace.so
( timeout 100 python scripts/corr_trace.py 2 1024 38 63 1; timeout 100 python scripts/corr_trace.py 2 2048 38 63 1 ) 2>&1 | grep -v Warn | cut -c1-200 | tee gpurun_out/r02_c50_corr_trace.txt
