#!/usr/bin/env python
"""One decoder-sized convolution through zvx_test_conv (PRO_F16 operand, 1056 -> 1056, k = 3, the bench batch's 64 utterance
lengths): a single-kernel workload for ncu and for checking the CTA-pair path against the plain one.
usage: python tools/conv_bench.py [cin cout k]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from zvxload import zvx
from zerovox_cpp_b200 import capi

cin, cout, k = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (1056, 1056, 3)
gguf = zvx.synth.write_model(zvx.synth.default_model_path())
ctx = capi.Context.from_gguf(gguf, device=0)
rows = [int(x) for x in zvx.synth.batch_lengths(64, seed=11)]
R = sum(rows)
rng = np.random.default_rng(1)
x = rng.standard_normal((R, cin)).astype(np.float16)
w = (rng.standard_normal((cout, cin, k)) / np.sqrt(cin * k)).astype(np.float16)
b = (rng.standard_normal(cout) * 0.1).astype(np.float32)
outs = []
for rep in range(3):
    t0 = time.perf_counter()
    outs.append(ctx.test_conv(rows, x, w, bias=b, pad=(k - 1) // 2, pro_mode=0))
    print("call", rep, "%.1f ms (host, copies included)" % ((time.perf_counter() - t0) * 1e3))
val = ctx.test_conv(rows, x, w, bias=b, pad=(k - 1) // 2, pro_mode=0, validation=True)
print("max |tensor-core path - validation kernel| =", float(np.abs(outs[-1] - val).max()), " repeatable:", bool(np.array_equal(outs[0], outs[-1])))
