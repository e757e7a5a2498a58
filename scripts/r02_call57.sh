#!/bin/bash
mkdir -p gpurun_out


timeout 200 python scripts/psroi_bwd_bench.py > gpurun_out/r02_c57_bwd.jsonl 2> gpurun_out/r02_c57_bwd.err
cat gpurun_out/r02_c57_bwd.jsonl | cut -c1-200; tail -3 gpurun_out/r02_c57_bwd.err
