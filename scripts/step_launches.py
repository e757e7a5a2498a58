#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum,... --csv` launch list of `bench.py --steps 1`:
kernels of ONE timed step (from one stem_pack_input launch to the next), grouped by name."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
idx = {n: hdr.index(n) for n in ("ID", "Kernel Name", "Metric Name", "Metric Value")}
per = collections.OrderedDict()
for r in rows[1:]:
    if not r[idx["ID"]].isdigit():
        continue
    k = int(r[idx["ID"]])
    d = per.setdefault(k, {"name": r[idx["Kernel Name"]]})
    try:
        d[r[idx["Metric Name"]]] = float(r[idx["Metric Value"]].replace(",", ""))
    except ValueError:
        pass
launches = list(per.values())
starts = [i for i, d in enumerate(launches) if "stem_pack_input" in d["name"]]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 3          # which step of the run (0-based)
a, b = starts[which], starts[which + 1]
step = launches[a:b]
step = [d for d in step if "CUDAFunctorOnSelf_add" not in d["name"] or d.get("gpu__time_duration.sum", 0) < 50000]   # drop the L2 flush
tot = sum(d.get("gpu__time_duration.sum", 0) for d in step) / 1e6
print("# one step = launches %d..%d: %d launches, %.3f ms summed kernel time (ncu: cold-cache, serialised -- compare SHARES)" % (a, b, len(step), tot))
g = collections.OrderedDict()
for d in step:
    n = d["name"]
    n = n[n.find("conv_igemm"):][:40] if "conv_igemm" in n else n[-70:]
    e = g.setdefault(n, [0, 0.0, 0.0, 0.0, 0.0])
    e[0] += 1
    e[1] += d.get("gpu__time_duration.sum", 0) / 1e6
    e[2] += d.get("dram__bytes_read.sum", 0) / 1e6
    e[3] += d.get("dram__bytes_write.sum", 0) / 1e6
    e[4] += d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0) * d.get("gpu__time_duration.sum", 0) / 1e6
for n, e in sorted(g.items(), key=lambda kv: -kv[1][1]):
    print("%-72s n=%4d  %8.3f ms %5.1f%%  dram rd %8.1f MB wr %8.1f MB  tensor-pipe active %4.1f%%" % (n, e[0], e[1], 100 * e[1] / tot, e[2], e[3], e[4] / e[1] if e[1] else 0))

if len(sys.argv) > 3:      # also write the traffic summary bench.py reads (profiles/rNN_traffic.json)
    import json
    conv = [d for d in step if "conv_igemm" in d["name"]]
    t_ns = sum(d.get("gpu__time_duration.sum", 0) for d in conv)
    out = {"source": "ncu launch list of `bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-train` (cold-cache, serialised), "
                     "summarised by scripts/step_launches.py",
           "conv_igemm": {
               "launches_per_step": len(conv),
               "dram_bytes_per_launch": sum(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0) for d in conv) / max(1, len(conv)),
               "dram_read_bytes_per_step": sum(d.get("dram__bytes_read.sum", 0) for d in conv),
               "dram_write_bytes_per_step": sum(d.get("dram__bytes_write.sum", 0) for d in conv),
               "ncu_time_ms_per_step": t_ns / 1e6,
               "share_of_step_kernel_time": t_ns / 1e6 / tot,
               "tensor_pipe_active_pct_time_weighted": sum(d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0) *
                                                           d.get("gpu__time_duration.sum", 0) for d in conv) / max(1.0, t_ns)}}
    try:
        prev = json.load(open(sys.argv[3]))
        for k, v in prev.items():
            out.setdefault(k, v)
    except (OSError, ValueError):
        pass
    json.dump(out, open(sys.argv[3], "w"), indent=1)
