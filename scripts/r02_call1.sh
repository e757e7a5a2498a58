#!/bin/bash
# round 2, GPU call 1: full GPU test suite (incl. the new Res-101 600x1000 parity test), the PSRoI backward experiment,
# the bench line and the training line
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s 2>&1 | tail -60 > gpurun_out/r02_c1_pytest.log
echo "pytest rc=$?" >> gpurun_out/r02_c1_pytest.log
D2T_TEST_EXPERIMENTS=1 timeout 300 python -m pytest tests/test_ops_gpu.py -q -k integer_tables_experiment 2>&1 | tail -15 > gpurun_out/r02_c1_psroi_bwd_test.log
timeout 300 python scripts/psroi_bwd_try.py > gpurun_out/r02_c1_psroi_bwd_try.jsonl 2> gpurun_out/r02_c1_psroi_bwd_try.err
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_c1_bench.json 2> gpurun_out/r02_c1_bench.err
echo "bench rc=$?" >> gpurun_out/r02_c1_bench.err
timeout 600 python bench.py --train --steps 3 --warmup 3 > gpurun_out/r02_c1_train.json 2> gpurun_out/r02_c1_train.err
echo "train rc=$?" >> gpurun_out/r02_c1_train.err
tail -5 gpurun_out/r02_c1_pytest.log
