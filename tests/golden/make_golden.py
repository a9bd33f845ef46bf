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

  regulator_*.npz     outputs of the UNMODIFIED reference program (oracle/_ref/zvfull_native = zerovox.cpp +
                      fs2encoder.cpp + ... compiled where they lie, oracle/ref_full_driver.cpp) on the random-init
                      GGUF that also carries the FastSpeech2 tensors (synth.write_model(with_fs2=True)):
                        regulator_default   the sentence hard-coded in ZeroVOXModel::eval (zerovox.cpp:204-314; recorded
                                            from the reference's own call, never copied from the source), full pipeline:
                                            src/puncts/style, features [120,528], log-durations, the frame count
                                            FS2Encoder::eval returns, the expanded hidden_state (stored as frame -> phoneme
                                            row map, verified bit-exact here), mel [1500,80], wav [450000]
                        regulator_random    a random phoneme sequence (seed 3), encoder + regulator only
                        regulator_capped    the default sentence with the duration predictor's bias raised to 3.0
                                            (another GGUF): the expansion hits max_seq_len -- fs2encoder.cpp:636-640

Inputs of ref_L*.npz are NOT stored: tests regenerate them from the seeds.
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


def regulator_fixture(name, res, with_audio):
    feat, hidden, frames = res["feat"], res["hidden"], int(res["frames"])
    # hidden_state is rows of `features` repeated + a zero tail: store the row map and check it is exact
    fmap = np.full(hidden.shape[0], -1, np.int16)
    i = 0
    for f in range(frames):
        while not np.array_equal(hidden[f].view(np.uint32), feat[i].view(np.uint32)):
            i += 1
        fmap[f] = i
    assert not hidden[frames:].any()
    rebuilt = np.zeros_like(hidden)
    rebuilt[:frames] = feat[fmap[:frames]]
    assert np.array_equal(rebuilt.view(np.uint32), hidden.view(np.uint32))
    d = dict(src=res["src"], puncts=res["puncts"], style=res["style"], feat=feat, logdur=res["logdur"],
             frames=np.int32(frames), frame_src=fmap, max_seq_len=np.int32(hidden.shape[0]))
    if with_audio:
        d.update(mel=res["mel"], wav=res["wav"])
    np.savez_compressed(os.path.join(HERE, f"regulator_{name}.npz"), **d)
    print(name, "frames", frames, "written")


def make_regulator():
    full = zvx.synth.write_model(zvx.synth.default_model_path(with_fs2=True), with_fs2=True)
    exe = os.path.join(ROOT, "oracle", "_ref", "zvfull_native")
    regulator_fixture("default", refrun.run_full(full, stages="full", binary=exe), True)
    rng = np.random.Generator(np.random.PCG64(3))
    src = rng.integers(1, zvx.synth.NUM_PHONEMES + 1, zvx.synth.MAX_N_PHONEMES).astype(np.int32)
    pun = rng.integers(0, 4, zvx.synth.MAX_N_PHONEMES).astype(np.int32)
    sty = (0.05 * rng.standard_normal(zvx.synth.DIM)).astype(np.float32)
    regulator_fixture("random", refrun.run_full(full, src, pun, sty, stages="enc", binary=exe), False)
    # same tensors except a larger duration bias: the expansion runs into max_seq_len
    from zerovox_cpp_b200.gguf_io import write_gguf
    t = zvx.synth.make_tensors()
    t.update(zvx.synth.make_fs2_tensors())
    t["_pe._var_adapt.duration_predictor.linear_layer.b"] = np.full((1,), 3.0, np.float32)
    capped = os.path.join(os.path.dirname(full), "zerovox-random-full-longdur.gguf")
    write_gguf(capped, zvx.synth.KV, t)
    res = refrun.run_full(capped, stages="enc", binary=exe)
    assert res["frames"] == zvx.synth.MAX_SEQ_LEN
    regulator_fixture("capped", res, False)
    os.unlink(capped)


def main():
    if "regulator" in sys.argv[1:]:
        make_regulator()
        return
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
    make_regulator()


if __name__ == "__main__":
    main()
