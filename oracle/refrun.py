"""oracle/refrun.py -- TEST INFRASTRUCTURE, not product code.

Runs the compiled, unmodified reference (oracle/_ref/zvref_*, built by oracle/Makefile
from /root/reference/src/{stylettsdec,hifigan,utils}.cpp + vendored ggml) in a fresh
process per call (one process per L: SURVEY.md 8c hazards H1/H2) and returns its
mel / wav as numpy arrays plus the timing line it prints.
"""
from __future__ import annotations

import json
import os
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref")
_picked = None
_picked_full = None


def ref_binary():
    """First of zvref_native / zvref_v3 that exists and executes on this host (the
    native build may hit SIGILL on a GPU-box CPU that differs from the build host)."""
    global _picked
    if _picked is not None:
        return _picked or None
    for name in ("zvref_native", "zvref_v3"):
        p = os.path.join(_REF, name)
        if not os.path.exists(p):
            continue
        try:
            r = subprocess.run([p], capture_output=True, timeout=30)
        except Exception:
            continue
        if r.returncode == 2:      # usage message: the binary executes
            _picked = p
            return p
    _picked = ""
    return None


def available() -> bool:
    return ref_binary() is not None


def run(gguf_path, L, enc=None, style=None, mel_in=None, stage="both", threads=0, reps=1,
        want_output=True, binary=None):
    """stage: both | dec | voc. Returns dict(mel=..., wav=..., timing=...)."""
    exe = binary or ref_binary()
    if exe is None:
        raise RuntimeError("oracle/_ref is not built (run `make -C oracle` where /root/reference exists)")
    threads = threads or os.cpu_count() or 1
    with tempfile.TemporaryDirectory(prefix="zvref_") as td:
        enc_p = sty_p = mel_p = "-"
        if stage in ("both", "dec"):
            enc_p = os.path.join(td, "enc.f32")
            sty_p = os.path.join(td, "sty.f32")
            np.ascontiguousarray(enc, dtype=np.float32).tofile(enc_p)
            np.ascontiguousarray(style, dtype=np.float32).tofile(sty_p)
        else:
            mel_p = os.path.join(td, "melin.f32")
            np.ascontiguousarray(mel_in, dtype=np.float32).tofile(mel_p)
        out = os.path.join(td, "out") if want_output else "-"
        r = subprocess.run([exe, gguf_path, str(int(L)), enc_p, sty_p, mel_p, out, str(threads), str(reps), stage],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"reference run failed ({r.returncode}): {r.stderr[-2000:]}")
        timing = None
        for line in r.stderr.splitlines():
            line = line.strip()
            if line.startswith("{"):
                timing = json.loads(line)
        res = {"timing": timing, "binary": os.path.basename(exe)}
        if want_output:
            if stage in ("both", "dec"):
                res["mel"] = np.fromfile(out + ".mel.f32", dtype=np.float32).reshape(L, -1)
            if stage in ("both", "voc"):
                res["wav"] = np.fromfile(out + ".wav.f32", dtype=np.float32)
        return res


def full_binary():
    """First of zvfull_native / zvfull_v3 (the whole unmodified program: FastSpeech2 encoder + length regulator +
    decoder + vocoder, oracle/ref_full_driver.cpp) that executes on this host."""
    global _picked_full
    if _picked_full is not None:
        return _picked_full or None
    for name in ("zvfull_native", "zvfull_v3"):
        p = os.path.join(_REF, name)
        if not os.path.exists(p):
            continue
        try:
            r = subprocess.run([p], capture_output=True, timeout=30)
        except Exception:
            continue
        if r.returncode == 2:
            _picked_full = p
            return p
    _picked_full = ""
    return None


def full_available() -> bool:
    return full_binary() is not None


def run_full(gguf_path, src=None, puncts=None, style=None, n_phonemes=None, stages="enc", threads=0, reps=1,
             want_output=True, binary=None):
    """Run ZeroVOXModel (reference, unmodified).  src/puncts/style None -> the sentence hard-coded in
    ZeroVOXModel::eval.  Returns dict(frames, src, puncts, style, feat [120,emb], logdur [120], hidden [max_seq_len,emb],
    mel, wav (stages == "full"), timing)."""
    exe = binary or full_binary()
    if exe is None:
        raise RuntimeError("oracle/_ref/zvfull_* is not built (run `make -C oracle` where /root/reference exists)")
    threads = threads or os.cpu_count() or 1
    with tempfile.TemporaryDirectory(prefix="zvfull_") as td:
        inp = "default"
        if src is not None:
            import struct
            inp = os.path.join(td, "in.bin")
            src = np.ascontiguousarray(src, dtype=np.int32)
            with open(inp, "wb") as f:
                f.write(struct.pack("<i", int(n_phonemes if n_phonemes is not None else len(src))))
                f.write(src.tobytes())
                f.write(np.ascontiguousarray(puncts, dtype=np.int32).tobytes())
                f.write(np.ascontiguousarray(style, dtype=np.float32).tobytes())
        out = os.path.join(td, "o") if want_output else "-"
        r = subprocess.run([exe, gguf_path, inp, out, str(threads), str(reps), stages], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"reference run failed ({r.returncode}): {r.stderr[-2000:]}")
        timing = None
        for line in r.stderr.splitlines():
            line = line.strip()
            if line.startswith("{"):
                timing = json.loads(line)
        res = {"timing": timing, "frames": timing["frames"], "binary": os.path.basename(exe)}
        if want_output:
            emb, T = timing["emb"], timing["max_seq_len"]
            res["src"] = np.fromfile(out + ".src.i32", dtype=np.int32)
            res["puncts"] = np.fromfile(out + ".puncts.i32", dtype=np.int32)
            res["style"] = np.fromfile(out + ".style.f32", dtype=np.float32)
            res["feat"] = np.fromfile(out + ".feat.f32", dtype=np.float32).reshape(-1, emb)
            res["logdur"] = np.fromfile(out + ".logdur.f32", dtype=np.float32)
            res["hidden"] = np.fromfile(out + ".hidden.f32", dtype=np.float32).reshape(T, emb)
            if stages == "full":
                res["mel"] = np.fromfile(out + ".mel.f32", dtype=np.float32).reshape(T, -1)
                res["wav"] = np.fromfile(out + ".wav.f32", dtype=np.float32)
        return res
