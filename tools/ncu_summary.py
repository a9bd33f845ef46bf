#!/usr/bin/env python
"""Text summary of an .ncu-rep (one launch): the metrics DESIGN.md / bench.py quote + the stall samples by code region.
usage: python tools/ncu_summary.py report.ncu-rep [launch index]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units, r = rows[0], rows[1], rows[2 + idx]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_op_read_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "sm__inst_executed.sum"]
for w in want:
    for i, n in enumerate(h):
        if n == w or n.endswith("." + w):
            print(f"{w:85s} {units[i]:16s} {r[i]}")
            break
print("(two tensor-pipe counters: sm__pipe_tensor_cycles_active = the metric B200_PROFILING.md greps and DESIGN.md quotes; the TriageCompute")
print(" *_realtime variant counts differently -- it reads lower for the same launch, 41 % vs 50 % in round 1's capture of the fused kernel)")
print()
out = subprocess.run([sys.executable, __file__.replace("ncu_summary.py", "ncu_regions.py"), rep, str(idx)], capture_output=True, text=True).stdout
print("\n".join(l for l in out.splitlines() if not l.startswith("name=")))
