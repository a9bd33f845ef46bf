#!/usr/bin/env python
"""Measurements for the BASELINE.json configs that bench.py's headline line does not cover
(run on the GPU box; prints one JSON line per config):
  configs[2]  long-form: 60 s utterance (L = 4800), decoder whole-sequence + vocoder streamed in
              overlapping 256-frame chunks (+-20-frame halo), host buffers in and out
  configs[4]  single-stream latency: one utterance at L = 400 (5 s) and at L = 1500 (the reference's
              shipped max_seq_len incl. zero tail semantics), p50 / p99 over N runs through the
              host-pointer C ABI (H2D + compute + D2H), i.e. StyleTTSDecoder::eval + HiFiGAN::eval
  configs[4] as BASELINE.json words it: the reference's hard-coded sentence through host/_build/zvx_model -- FastSpeech2
              encoder on the host (the reference's own fs2encoder.cpp + ggml CPU), length regulator + decoder + vocoder on
              the B200 -- p50 / p99 of encoder, GPU call and total over N runs, in the reference's default mode (1500 frames
              with the zero tail) and with only the valid frames synthesised."""
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
    # ---- configs[4] including the FastSpeech2 encoder (host) ----
    import struct
    import subprocess
    import tempfile
    exe = os.path.join(ROOT, "zerovox.cpp_b200", "host", "_build", "zvx_model")
    gold = os.path.join(ROOT, "tests", "golden", "regulator_default.npz")
    if os.path.exists(exe) and os.path.exists(gold):
        g = np.load(gold)
        full = zvx.synth.write_model(zvx.synth.default_model_path(with_fs2=True), with_fs2=True)
        with tempfile.TemporaryDirectory() as td:
            sp = os.path.join(td, "s0.bin")
            with open(sp, "wb") as f:
                f.write(struct.pack("<i", 120))
                f.write(np.ascontiguousarray(g["src"], np.int32).tobytes())
                f.write(np.ascontiguousarray(g["puncts"], np.int32).tobytes())
                f.write(np.ascontiguousarray(g["style"], np.float32).tobytes())
            for mode in (["--reference-default"], []):
                r = subprocess.run([exe, full, os.path.join(td, "o")] + mode + ["--bench", str(runs), sp], capture_output=True, text=True)
                if r.returncode != 0:
                    print(json.dumps({"config": "configs[4] with encoder", "error": r.stderr[-500:]}))
                    continue
                j = json.loads(r.stdout.strip().splitlines()[-1])
                j["config"] = ("configs[4] e2e latency: FS2 encoder on the host (reference's fs2encoder.cpp, ggml CPU) + length regulator, "
                               "decoder, vocoder on 1xB200, the reference's default sentence (zerovox.cpp:204-314), " + j["mode"])
                print(json.dumps(j))
    else:
        print(json.dumps({"config": "configs[4] with encoder", "skipped": "host/_build/zvx_model or the regulator golden missing"}))


if __name__ == "__main__":
    main()
