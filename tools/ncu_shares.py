#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launch count, total time and share.

  python tools/ncu_shares.py gpurun_out/launches.csv [title]  > profiles/rNN_ncu_launch_shares.txt"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.defaultdict(float)
cnt = collections.Counter()
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1e-3)
    name = re.sub(r"\(.*$", "", r[ki]).replace("void ", "")[:96]
    tot[name] += v * scale
    cnt[name] += 1
total = sum(tot.values())
print("# %s" % (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]))
print("# per-launch times under ncu are cold-cache and serialised: compare SHARES.  total %.1f us in %d launches" % (
    total, sum(cnt.values())))
for name, t in sorted(tot.items(), key=lambda kv: -kv[1])[:40]:
    print("%-98s n=%5d %10.1f us %5.1f%%" % (name, cnt[name], t, 100.0 * t / total))
