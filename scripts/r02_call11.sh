#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_engine_gpu.py tests/test_train_conv_gpu.py -q 2>&1 | tail -5 > gpurun_out/r02_c11_tests.log
NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_c11_bench_n2.json 2> gpurun_out/r02_c11_bench_n2.err
echo "rc=$?" >> gpurun_out/r02_c11_bench_n2.err
tail -n 3 gpurun_out/r02_c11_tests.log; grep -c NCCL gpurun_out/r02_c11_bench_n2.err; tail -n 3 gpurun_out/r02_c11_bench_n2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_c11_bench_n2.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["n_gpus"])
print(json.dumps(d["train"])[:1500])
PY
