#!/bin/bash
mkdir -p gpurun_out
make -s -C pytorch-detect-to-track_b200/csrc trace > gpurun_out/r02_c33_make.log 2>&1
export D2T_B200_LIB=$PWD/pytorch-detect-to-track_b200/d2t_b200/libd2t_b200_trace.so
timeout 120 python scripts/chain_trace.py 2 2>&1 | grep -v Warn | tail -8 | cut -c1-900
