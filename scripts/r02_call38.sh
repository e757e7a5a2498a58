#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_c38_bench_n2.json 2> gpurun_out/r02_c38_bench_n2.err
echo "exit $?"
tail -c 500 gpurun_out/r02_c38_bench_n2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_c38_bench_n2.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["n_gpus"], d["e2e"]["value"], d["parity"]["ok"])
print({k: d["train"][k] for k in ("ms_per_step", "value", "nranks", "allreduce_ms", "allreduce_exposed_ms", "overlap_frac", "engine_forward_ms", "engine_backward_ms", "heads_losses_optimizer_ms", "loss_finite")})
PY
