#!/usr/bin/env python
"""Tiny end-to-end run (decoder + vocoder, ragged batch incl. edge windows) for compute-sanitizer."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from zvxload import zvx
from zerovox_cpp_b200 import capi
gguf = zvx.synth.write_model(zvx.synth.default_model_path())
ctx = capi.Context.from_gguf(gguf, device=0)
ins = [zvx.synth.make_inputs(L, seed=3 + i) for i, L in enumerate((5, 48, 17))]
mels, wavs = ctx.synth_batch([e for e, _ in ins], [s for _, s in ins])
print("ok", [float(np.abs(w).max()) for w in wavs])
ctx.close()
