#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel.  usage: launch_summary.py launches.csv out.csv "comment" """
import csv, sys, collections, re
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
tot = collections.defaultdict(float); n = collections.Counter()
for r in rows[1:]:
    if len(r) <= vi: continue
    v = float(r[vi].replace(",", "")); u = r[ui]
    us = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3 if u in ("ms", "msecond") else v)
    k = re.sub(r"\(.*", "", r[ki]); tot[k] += us; n[k] += 1
s = sum(tot.values())
with open(sys.argv[2], "w") as f:
    f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none); cold-cache, serialised: compare shares\n# %s\nkernel,launches,total_us,share\n" % (sys.argv[3] if len(sys.argv) > 3 else ""))
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        f.write("%s,%d,%.1f,%.4f\n" % (k, n[k], v, v / s))
print(open(sys.argv[2]).read())
