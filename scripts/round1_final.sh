#!/bin/bash
# The round's closing GPU call: tests, the bench line, ncu captures of the new PSRoI kernel, the step's launch list.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/fin_status.txt
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/fin_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/fin_status.txt
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/fin_bench.json 2> gpurun_out/fin_bench.err
echo "bench exit $?" >> gpurun_out/fin_status.txt
timeout 200 ncu --set full --import-source on --clock-control none -k regex:psroi_fwd -c 2 -f -o gpurun_out/r01_psroi_mc_b2 python scripts/psroi_one.py 2 > gpurun_out/fin_ncu_b2.log 2>&1
echo "ncu psroi b2 exit $?" >> gpurun_out/fin_status.txt
timeout 200 ncu --set full --import-source on --clock-control none -k regex:psroi_fwd -c 2 -f -o gpurun_out/r01_psroi_mc_b8 python scripts/psroi_one.py 8 > gpurun_out/fin_ncu_b8.log 2>&1
echo "ncu psroi b8 exit $?" >> gpurun_out/fin_status.txt
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -c 1600 --csv --log-file gpurun_out/fin_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/fin_launches_bench.log 2>&1
echo "ncu launch list exit $?" >> gpurun_out/fin_status.txt
cat gpurun_out/fin_status.txt
tail -n 3 gpurun_out/fin_pytest.log
