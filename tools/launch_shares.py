"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) into per-kernel totals and shares.
Usage: python tools/launch_shares.py gpurun_out/launches.csv "<title line>" > profiles/rN_launch_shares.txt"""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[1:]:
    if len(r) <= iv:
        continue
    v = float(r[iv].replace(",", ""))
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
    tot[r[ik]] += v * scale
    cnt[r[ik]] += 1
allms = sum(tot.values())
print("# " + (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]))
print(f"{'kernel':76s} {'n':>4s} {'total ms':>10s} {'share':>7s}")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{k[:76]:76s} {cnt[k]:4d} {v:10.3f} {100 * v / allms:6.1f}%")
