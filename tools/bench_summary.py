#!/usr/bin/env python
"""Print a one-screen summary of a bench.py JSON line read from stdin."""
import json, sys
d = json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1])
print(f"value={d['value']:.0f} {d['unit']}  ms/step={d['ms_per_step']:.2f}  e2e={d['e2e']['value']:.0f}  mrf={d['roofline']['achieved']:.0f} TFLOP/s ({100*d['roofline']['frac']:.1f}%)")
print({k: (v["ms_per_step"], v["tflops"]) for k, v in d["kernel_breakdown"].items()})
