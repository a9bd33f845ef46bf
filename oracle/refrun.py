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
