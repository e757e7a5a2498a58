#!/bin/bash
mkdir -p gpurun_out
D2T_B200_LIB=$PWD/pytorch-detect-to-track_b200/d2t_b200/libd2t_b200_trace.so timeout 200 python scripts/psroi_fwd_trace.py > gpurun_out/r02_c62_trace.txt 2>&1
tail -3 gpurun_out/r02_c62_trace.txt
