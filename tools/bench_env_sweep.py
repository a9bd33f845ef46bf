#!/usr/bin/env python
"""Developer tool: run bench.py under several values of one environment switch of libzvx.so and print the
headline numbers side by side.  usage: python tools/bench_env_sweep.py ZVX_E2E_CHUNKS 2 3 4 [--steps 5]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
args = sys.argv[1:]
steps = "5"
if "--steps" in args:
    i = args.index("--steps")
    steps = args[i + 1]
    del args[i:i + 2]
key, values = args[0], args[1:]
for v in values:
    env = dict(os.environ, **{key: v})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--no-cpu-baseline", "--steps", steps], env=env,
                         capture_output=True, text=True).stdout
    line = [l for l in out.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    print(f"{key}={v}: value={d['value']:.0f} ms/step={d['ms_per_step']:.3f} e2e={d['e2e']['value']:.0f} "
          f"e2e_pcm16={d['e2e_pcm16']['value']:.0f} e2e_regulated={d['e2e_regulated_pcm16']['value']:.0f}", flush=True)
