# A/B experiments on one box with the trace build (make -C pytorch-detect-to-track_b200/csrc trace)
export D2T_B200_LIB=$PWD/pytorch-detect-to-track_b200/d2t_b200/libd2t_b200_trace.so
for e in 1 0; do
  for cfg in "4 256 38 63 1024 1 1 0 1 16 res" "4 64 150 250 256 1 1 0 1 16" "4 128 75 125 512 1 1 0 1 16 res" "4 1024 38 63 256 1 1 0 1 16" "4 256 38 63 256 3 1 1 1 16"; do
    echo "EPI2=$e $(D2T_CONV_EPI2=$e timeout 120 python scripts/conv_trace.py $cfg 2>&1 | grep -E 'layer|epilogue0|mma' | sed -e 's/info.*grid.: [0-9]*}//' | tr '\n' ' ')"
  done
done
