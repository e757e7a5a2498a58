#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/r02_heads_profile.py 2>&1 | grep -v Warn > gpurun_out/r02_c47_train_profile.txt
head -50 gpurun_out/r02_c47_train_profile.txt | cut -c1-60,120-230
