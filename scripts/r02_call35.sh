#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r02_c35_bench.json 2> gpurun_out/r02_c35_bench.err
echo "bench exit $?"; tail -c 400 gpurun_out/r02_c35_bench.err
timeout 240 ncu --set full --import-source on --clock-control none -k regex:conv_igemm -s 2 -c 1 -f -o gpurun_out/r02_conv_ares_final python scripts/conv_prof1.py 4 256 38 63 1024 1 1 0 1 16 res > gpurun_out/r02_c35_ncu.log 2>&1
echo "ncu exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_c35_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["conv_ms_per_step"], d["e2e"]["value"], d["parity"]["ok"], d["train"]["ms_per_step"])
for k, v in d["ops"].items():
    if k.startswith("roi_") or k.startswith("psroi_vote") or k == "corr_conv4":
        print(k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items() if a not in ("note", "kernel", "shape")})
PY
