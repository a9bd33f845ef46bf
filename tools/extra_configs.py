#!/usr/bin/env python
"""Measurements for the BASELINE.json configs that bench.py's headline line does not cover
(run on the GPU box; prints one JSON line per config):
  configs[2]  long-form: 60 s utterance (L = 4800), decoder whole-sequence + vocoder streamed in
              overlapping 256-frame chunks (+-20-frame halo), host buffers in and out
  configs[4]  single-stream latency: one utterance at L = 400 (5 s) and at L = 1500 (the reference's
              shipped max_seq_len incl. zero tail semantics), p50 / p99 over N runs through the
              host-pointer C ABI (H2D + compute + D2H), i.e. StyleTTSDecoder::eval + HiFiGAN::eval
(the FS2 encoder of configs[4] stays on the reference's ggml CPU path and is not timed here)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zvxload import zvx  # noqa: E402
from zerovox_cpp_b200 import capi  # noqa: E402


def main():
    runs = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    gguf = zvx.synth.write_model(zvx.synth.default_model_path())
    ctx = capi.Context.from_gguf(gguf, device=0)
    # ---- long-form ----
    L = 4800
    enc, sty = zvx.synth.make_inputs(L, seed=5)
    for _ in range(2):
        mel = ctx.decode(enc, sty)
        ctx.vocode_chunked(mel, 256, 20)
    ts = []
    for _ in range(10):
        t0 = time.perf_counter()
        mel = ctx.decode(enc, sty)
        wav = ctx.vocode_chunked(mel, 256, 20)
        ts.append(time.perf_counter() - t0)
    first = []
    for _ in range(10):
        t0 = time.perf_counter()
        stamp = []
        ctx.vocode_chunked(mel, 256, 20, on_chunk=lambda f, n, w: stamp.append(time.perf_counter() - t0))
        first.append(stamp[0])
    tw = []
    for _ in range(10):
        t0 = time.perf_counter()
        mel = ctx.decode(enc, sty)
        wav2 = ctx.vocode(mel)
        tw.append(time.perf_counter() - t0)
    print(json.dumps({"config": "configs[2] long-form 60 s utterance, 1xB200, host buffers", "L": L,
                      "chunked_256+20_ms_median": 1e3 * float(np.median(ts)), "whole_sequence_ms_median": 1e3 * float(np.median(tw)),
                      "first_chunk_ms_median": 1e3 * float(np.median(first)), "audio_s_per_s_chunked": 60.0 / float(np.median(ts)), "audio_s_per_s_whole": 60.0 / float(np.median(tw)),
                      "max_abs_chunked_vs_whole": float(np.abs(wav - wav2).max())}))
    # ---- latency ----
    for L in (400, 1500):
        enc, sty = zvx.synth.make_inputs(L, seed=7)
        for _ in range(5):
            ctx.vocode(ctx.decode(enc, sty))
        lat = []
        for _ in range(runs):
            t0 = time.perf_counter()
            ctx.vocode(ctx.decode(enc, sty))
            lat.append(1e3 * (time.perf_counter() - t0))
        lat = np.sort(np.array(lat))
        print(json.dumps({"config": "configs[4] single-stream latency, decoder.eval + vocoder.eval via host-pointer C ABI", "L": L,
                          "audio_s": L / 80.0, "runs": runs, "p50_ms": float(lat[len(lat) // 2]), "p99_ms": float(lat[int(len(lat) * 0.99) - 1]),
                          "min_ms": float(lat[0]), "rtf_inverse_p50": (L / 80.0) / (float(lat[len(lat) // 2]) / 1e3)}))
    ctx.close()


if __name__ == "__main__":
    main()
