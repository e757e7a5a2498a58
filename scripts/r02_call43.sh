#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v Warn | tail -3
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_c43_ref.json 2> gpurun_out/r02_c43_ref.err
echo "ref exit $?"; cat gpurun_out/r02_c43_ref.json | cut -c1-700
