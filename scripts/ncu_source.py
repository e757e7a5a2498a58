#!/usr/bin/env python
"""Per-CUDA-line instruction and stall-sample totals from an .ncu-rep.
usage: scripts_ncu_source.py file.ncu-rep kernel_regex [top_n]"""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--print-source", "cuda,sass", "--csv", "-k",
                      "regex:" + sys.argv[2]], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
fname, data, hdr = "", [], None
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        fname = r[1].split('/')[-1]
    elif len(r) > 6 and r[0] == 'Line No':
        hdr = r
    elif hdr and len(r) > 7 and r[0].isdigit():
        try:
            data.append((int(r[hdr.index('Instructions Executed')]), int(r[hdr.index('# Samples')]), fname, r[0], r[1].strip()[:100]))
        except ValueError:
            pass
ti, ts = sum(d[0] for d in data), sum(d[1] for d in data)
print("total warp-instructions %d, samples %d" % (ti, ts))
for n, s, f, l, src in sorted(data, key=lambda d: -d[1])[:top]:
    print("%5.1f%% smp %5.1f%% inst  %s:%s  %s" % (100.0 * s / max(ts, 1), 100.0 * n / max(ti, 1), f, l, src))
