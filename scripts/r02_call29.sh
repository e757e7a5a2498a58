#!/bin/bash
mkdir -p gpurun_out
timeout 240 ncu --set full --import-source on --clock-control none -k regex:conv_igemm -s 2 -c 1 -f -o gpurun_out/r02_conv_ares_c python scripts/conv_prof1.py 4 256 38 63 1024 1 1 0 1 16 res > gpurun_out/r02_c29_ncu.log 2>&1
echo "ncu exit $?"
ls -la gpurun_out/r02_conv_ares_c.ncu-rep
