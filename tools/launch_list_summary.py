#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X ...`).

    python tools/launch_list_summary.py gpurun_out/launches.csv > profiles/rNN_launches_summary.txt

Per-launch times under ncu are cold-cache and serialised: compare the SHARES with bench.py's kernel_breakdown, not the
absolute times."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hi]
ci = {n: i for i, n in enumerate(h)}
agg = collections.OrderedDict()
tot = 0.0
for r in rows[hi + 1:]:
    if len(r) < len(h):
        continue
    name = re.sub(r"^void\s+", "", r[ci["Kernel Name"]])
    name = re.sub(r"\(.*$", "", name).replace("zvx::", "").replace("<unnamed>::", "")
    v = float(r[ci["Metric Value"]].replace(",", ""))
    unit = r[ci["Metric Unit"]]
    us = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
    tot += us
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:58s} launches={n:4d} total_us={us:10.1f} share={100 * us / tot:5.1f}%")
print(f"{'all':58s} launches={sum(a[0] for a in agg.values()):4d} total_us={tot:10.1f}")
