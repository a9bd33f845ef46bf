#!/usr/bin/env python
"""Summarise an ncu report of the fused MRF kernel: per launch duration / tensor-pipe %, and the
warp-stall samples grouped into code regions delimited by marker instructions (source page)."""
import csv, subprocess, sys, io

rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else -1
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
ci = {n: i for i, n in enumerate(h)}
def col(name):
    for n in h:
        if name in n: return ci[n]
    return None
cols = [("Kernel Name", "name"), ("Grid Size", "grid"), ("gpu__time_duration.sum", "ms"),
        ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor%"),
        ("sm__inst_executed.sum", "inst"), ("sm__cycles_elapsed.max", "cycles")]
for r in rows[2:]:
    print("  ".join(f"{lab}={r[col(n)][:40]}" for n, lab in cols if col(n) is not None))
if which < 0: sys.exit(0)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi = hdr_idx[which]
end = hdr_idx[which + 1] - 1 if which + 1 < len(hdr_idx) else len(rows)
h = rows[hi]; body = rows[hi + 1:end]
ci = {n: i for i, n in enumerate(h)}
S, I, SRC = ci["# Samples"], ci["Instructions Executed"], ci["Source"]
tot = sum(int(r[S]) for r in body)
print("kernel", rows[hi - 1][1][:80], "total samples", tot, "instrs", len(body))
marks = [i for i, r in enumerate(body) if any(k in r[SRC] for k in ("STTM", "LDTM", "UTCHMMA", "UBLKCP", "BAR.SYNC", "SYNCS.ARRIVE", "UTCBAR", "EXIT", "TRYWAIT"))]
prev = 0
stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
for m in marks + [len(body)]:
    seg = body[prev:m]
    ssum = sum(int(r[S]) for r in seg)
    st = sorted([(sum(int(r[ci[n]]) for r in seg), n[6:]) for n in stall_cols], reverse=True)[:3]
    nxt = body[m][SRC].strip()[:60] if m < len(body) else "END"
    if ssum > tot * 0.004:
        print(f"[{prev:5d},{m:5d}) {ssum:6d} ({100.0*ssum/tot:4.1f}%) inst={sum(int(r[I]) for r in seg):9d} {st}  -> {nxt}")
    prev = m
