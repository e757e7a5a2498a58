"""Per-kernel counts of the Blackwell-specific SASS mnemonics in libd2t_b200.so (cuobjdump -sass):
UTCHMMA / UTCQMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA tensor load / store),
UBLKCP (1-D bulk copy), SYNCS (mbarrier).  usage: python scripts/sass_summary.py > profiles/rNN_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "pytorch-detect-to-track_b200", "d2t_b200", "libd2t_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
keys = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "FFMA", "HFMA2", "ATOM", "RED"]
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(anonymous namespace\)::", "", cur)
        cur = re.sub(r"\(.*", "", cur)[:110]
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        counts[cur]["_total"] += 1
        for k in keys:
            if op.startswith(k):
                counts[cur][k] += 1
arch = re.findall(r"arch = (sm_\w+)", out)
print("libd2t_b200.so: %d kernels, archs %s" % (len(counts), sorted(set(arch))))
print("%-112s %7s " % ("kernel", "instrs") + " ".join("%8s" % k for k in keys))
tot = collections.Counter()
for name, c in counts.items():
    print("%-112s %7d " % (name, c["_total"]) + " ".join("%8d" % c[k] for k in keys))
    tot.update(c)
print("%-112s %7d " % ("TOTAL", tot["_total"]) + " ".join("%8d" % tot[k] for k in keys))
