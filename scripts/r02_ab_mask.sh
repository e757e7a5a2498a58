#!/bin/bash
# A/B on ONE box: the conv forward chain with and without the backward-data mask code compiled into the epilogue
cd pytorch-detect-to-track_b200/csrc
mkdir -p build/ab
for f in *.cu; do nvcc -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -gencode arch=compute_100a,code=sm_100a -DD2T_AB_NO_MASK -c -o build/ab/$(basename $f .cu).o $f & done; wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../d2t_b200/libd2t_b200_ab.so build/ab/*.o -Xlinker --exclude-libs,ALL
