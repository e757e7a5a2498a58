#!/bin/bash
# Round-2 evidence run: the whole GPU suite, the default bench line, the step's ncu launch list (+ tensor-pipe activity),
# ncu --set full captures of representative conv layers and of the PSRoI kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/r02_fin_status.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_fin_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02_fin_status.txt
timeout 600 python bench.py > gpurun_out/r02_fin_bench.json 2> gpurun_out/r02_fin_bench.err
echo "bench exit $?" >> gpurun_out/r02_fin_status.txt
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -c 1700 --csv --log-file gpurun_out/r02_fin_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-train --no-graph > gpurun_out/r02_fin_launches_bench.log 2>&1
echo "ncu launch list exit $?" >> gpurun_out/r02_fin_status.txt
cat gpurun_out/r02_fin_status.txt
tail -n 3 gpurun_out/r02_fin_pytest.log
tail -c 600 gpurun_out/r02_fin_bench.err
timeout 300 python scripts/r02_heads_profile.py > gpurun_out/r02_fin_train_profile.txt 2>&1
echo "train profile exit $?" >> gpurun_out/r02_fin_status.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_fin_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/r02_fin_status.txt
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_fin_bench_reference.json 2> gpurun_out/r02_fin_bench_reference.err
echo "reference arm exit $?" >> gpurun_out/r02_fin_status.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"psroi_bwd_limb|psroi_bwd_amax|roi_align_bwd|det_finish" -c 8 -o gpurun_out/r02_fin_psroi_bwd python bench.py --ops-only > gpurun_out/r02_fin_psroi_bwd_ncu.log 2>&1
echo "ncu psroi bwd exit $?" >> gpurun_out/r02_fin_status.txt
tail -n 5 gpurun_out/r02_fin_status.txt
