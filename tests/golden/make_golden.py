"""Regenerates the committed golden fixtures of tests/golden/ (run where /root/reference exists).

  ref_L<L>.npz        outputs of the UNMODIFIED reference (oracle/_ref/zvref_native, built by
                      oracle/Makefile from /root/reference/src/{stylettsdec,hifigan,utils}.cpp + ggml)
                      on the deterministic random-init GGUF (synth.write_model, seed 1234) and
                      synth.make_inputs(L, seed 7):  mel [L,80], wav [L*300]; plus the same
                      reference rebuilt for another ISA (zvref_v3) as `wav_v3`/`mel_v3`, which
                      gives the reference's own self-noise floor.
  norm1d_example.npz  the reference's only known-answer fixture for this path,
                      /root/reference/utils/norm1dexample.json (PyTorch InstanceNorm1d(528, affine)
                      input/output 1x528x115, weight, bias), converted to float32 arrays.

Inputs are NOT stored: tests regenerate them from the seeds.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from zvxload import zvx  # noqa: E402
import refrun  # noqa: E402

LENGTHS = (48, 160, 400)


def main():
    gguf = zvx.synth.write_model(zvx.synth.default_model_path())
    native = os.path.join(ROOT, "oracle", "_ref", "zvref_native")
    v3 = os.path.join(ROOT, "oracle", "_ref", "zvref_v3")
    for L in LENGTHS:
        enc, sty = zvx.synth.make_inputs(L)
        a = refrun.run(gguf, L, enc, sty, binary=native)
        b = refrun.run(gguf, L, enc, sty, binary=v3)
        np.savez_compressed(os.path.join(HERE, f"ref_L{L}.npz"), mel=a["mel"], wav=a["wav"],
                            mel_v3=b["mel"], wav_v3=b["wav"])
        print(L, "written")
    src = "/root/reference/utils/norm1dexample.json"
    with open(src) as f:
        j = json.load(f)
    np.savez_compressed(os.path.join(HERE, "norm1d_example.npz"),
                        x_in=np.asarray(j["x_in"], np.float32), x_out=np.asarray(j["x_out"], np.float32),
                        weight=np.asarray(j["weight"], np.float32), bias=np.asarray(j["bias"], np.float32))
    print("norm1d written")


if __name__ == "__main__":
    main()
