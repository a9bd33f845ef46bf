#!/usr/bin/env python
"""A/B comparison of per-launch tables (tools/launch_table.py output) of two library builds measured in the
same GPU session: python tools/ab_compare.py A1.txt B1.txt [A2.txt B2.txt ...] -> ms per (kind, stage), min over repeats."""
import collections
import sys


def load(path):
    d = collections.OrderedDict()
    for line in open(path):
        f = line.split()
        if len(f) < 6 or not f[0].isdigit():
            continue
        key = f"{f[1]}:{f[2]}"
        d[key] = d.get(key, 0.0) + float(f[f.index("ms=") + 1])
    return d


files = sys.argv[1:]
A = [load(f) for f in files[0::2]]
B = [load(f) for f in files[1::2]]
keys = list(A[0].keys())
ta = tb = 0.0
for k in keys:
    a = min(x.get(k, 0.0) for x in A)
    b = min(x.get(k, 0.0) for x in B)
    ta += a
    tb += b
    print(f"{k:22s} A={a:8.4f} B={b:8.4f}  B/A={b / a if a else 0:6.3f}")
print(f"{'total':22s} A={ta:8.4f} B={tb:8.4f}  B/A={tb / ta:6.3f}")
