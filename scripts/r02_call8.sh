#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r02_c8_pytest.log
timeout 900 python scripts/r02_bwd_profile.py > gpurun_out/r02_c8_bwd_profile.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_c8_bench.json 2> gpurun_out/r02_c8_bench.err
echo "bench rc=$?" >> gpurun_out/r02_c8_bench.err
tail -n 25 gpurun_out/r02_c8_pytest.log; grep -v Warn gpurun_out/r02_c8_bwd_profile.log | head -12; tail -n 3 gpurun_out/r02_c8_bench.err
