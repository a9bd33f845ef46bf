#!/usr/bin/env python
"""Per-launch table of one step of the bench workload (kind, stage, algorithmic GFLOP, ms, TFLOP/s)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from zvxload import zvx
from zerovox_cpp_b200 import capi
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
gguf = zvx.synth.write_model(zvx.synth.default_model_path())
ctx = capi.Context.from_gguf(gguf, device=0)
lengths = zvx.synth.batch_lengths(B, seed=11) if len(sys.argv) < 3 else __import__("numpy").full(B, int(sys.argv[2]), dtype="int32")
F = int(lengths.sum())
Larr = (ctypes.c_int32 * B)(*[int(x) for x in lengths])
ctx.reserve(F, B)
d_enc = torch.randn(F, ctx.dim_in, device="cuda")
d_sty = 0.05 * torch.randn(B, ctx.style_dim, device="cuda")
d_wav = torch.empty(F * ctx.hop, device="cuda")
for _ in range(3):
    ctx.synth_batch_device(B, d_enc.data_ptr(), d_sty.data_ptr(), Larr, 0, d_wav.data_ptr(), sync=True)
ctx.profile_begin()
ctx.synth_batch_device(B, d_enc.data_ptr(), d_sty.data_ptr(), Larr, 0, d_wav.data_ptr(), sync=True)
recs = ctx.profile_end()
tot = sum(r[4] for r in recs)
for i, (kind, stage, flops, nbytes, ms) in enumerate(recs):
    tf = flops / (ms / 1e3) / 1e12 if ms > 0 and flops > 0 else 0.0
    print(f"{i:3d} {kind:14s} st={stage} gflop={flops/1e9:9.2f} ms={ms:7.4f} tflops={tf:7.1f}")
print("total kernel ms", tot)
