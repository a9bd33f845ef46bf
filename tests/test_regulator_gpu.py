"""GPU tests of the length regulator row: zvx_synth_batch_regulated (phoneme-rate features
expanded on the GPU) must equal zvx_synth_batch fed the matrix the restated host loop (fs2encoder.cpp:611-654)
produces -- bit-exact, the regulator is index work -- in both modes: reference-default (max_seq_len frames, zero
tail) and valid-frames-only.  Pinned against the unmodified reference program: tests/golden/regulator_*.npz hold what
ZeroVOXModel::eval / FS2Encoder::eval produced (hidden_state, frame count, mel, wav at max_seq_len = 1500)."""
import numpy as np
import pytest

import zv_oracle
from test_regulator_cpu import regulator_golden

pytestmark = pytest.mark.gpu


def test_regulator_matches_reference_program(ctx):
    """The three reference runs as ONE batch in reference-default mode: the decoder input the GPU expansion writes is
    bit-identical to the reference's hidden_state, frame counts are the reference's, and the default sentence's mel / wav
    (1500 frames, zero tail, statistics over the tail: SURVEY.md N2) meet the parity gates."""
    gs = [regulator_golden(n) for n in ("default", "random", "capped")]
    T = gs[0][3]
    valid, wavs, mels = ctx.synth_batch_regulated([g[0]["feat"] for g in gs], [g[0]["logdur"] for g in gs],
                                                  [g[0]["style"] for g in gs], T, True, want_mel=True)
    assert valid == [g[2] for g in gs]
    dim = gs[0][0]["feat"].shape[1]
    g0 = gs[0][0]
    assert zv_oracle.snr_db(g0["mel"], mels[0]) >= 55.0
    assert zv_oracle.snr_db(g0["wav"], wavs[0]) >= 60.0
    assert float(np.abs(g0["wav"] - wavs[0]).max()) <= 1e-3
    # decoder input: a single-lane call keeps all three utterances in one workspace
    for g, hidden, frames, _ in gs:
        v, _ = ctx.synth_batch_regulated([g["feat"]], [g["logdur"]], [g["style"]], T, True)
        assert v == [frames]
        x = ctx.debug_fetch("enc_in", T * dim).reshape(T, dim)
        assert np.array_equal(x.view(np.uint32), hidden.view(np.uint32))


def test_valid_frames_mode_matches_own_length_reference(ctx, gguf_path):
    """f1: synthesising only the valid frames equals the reference built for exactly that length (live reference run)."""
    import refrun
    g, hidden, frames, T = regulator_golden("default")
    valid, wavs, mels = ctx.synth_batch_regulated([g["feat"]], [g["logdur"]], [g["style"]], T, False, want_mel=True)
    assert valid == [frames] and wavs[0].size == frames * 300
    ref = refrun.run(gguf_path, frames, hidden[:frames], g["style"])
    assert zv_oracle.snr_db(ref["mel"], mels[0]) >= 55.0
    assert zv_oracle.snr_db(ref["wav"], wavs[0]) >= 60.0
    assert float(np.abs(ref["wav"] - wavs[0]).max()) <= 1e-3


def _utterances(zvx, n, seed, dim=528):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        P = int(rng.integers(3, 60))
        feats = rng.standard_normal((P, dim)).astype(np.float32)
        logd = rng.normal(1.3, 0.7, P).astype(np.float32)
        logd[rng.integers(0, P)] = -5.0                        # a phoneme of zero frames
        sty = (0.05 * rng.standard_normal(dim)).astype(np.float32)
        out.append((feats, logd, sty))
    return out


@pytest.mark.parametrize("pad_to_max,cap", [(True, 400), (False, 400), (False, 60), (True, 1000)])   # the last one is large enough for the two-lane pipelining
def test_regulated_synthesis_equals_host_expansion_bit_exact(ctx, zvx, pad_to_max, cap):
    utts = _utterances(zvx, 5, seed=11)
    valid, wavs = ctx.synth_batch_regulated([u[0] for u in utts], [u[1] for u in utts], [u[2] for u in utts], cap, pad_to_max)
    encs = []
    for (f, d, _), n in zip(utts, valid):
        x, n_ref = zv_oracle.length_regulate(f, d, cap)
        assert n == n_ref and 0 < n <= cap
        encs.append(x if pad_to_max else x[:n])
    _, ref = ctx.synth_batch(encs, [u[2] for u in utts], want_mel=False)
    for w, r in zip(wavs, ref):
        assert w.shape == r.shape and np.array_equal(w, r)


def test_regulated_pcm16_and_single_phoneme(ctx, zvx):
    rng = np.random.default_rng(5)
    f = rng.standard_normal((1, 528)).astype(np.float32)
    d = np.array([np.log(9.0)], np.float32)                    # one phoneme, 8 frames
    s = (0.05 * rng.standard_normal(528)).astype(np.float32)
    valid, pcm = ctx.synth_batch_regulated([f], [d], [s], 1500, False, pcm16=True)
    assert valid == [8] and pcm[0].dtype == np.int16 and pcm[0].size == 8 * 300
    x, _ = zv_oracle.length_regulate(f, d, 1500)
    _, ref = ctx.synth_batch([x[:8]], [s], want_mel=False)
    assert np.array_equal(pcm[0], zv_oracle.pcm16(ref[0]))


def test_regulated_rejects_all_zero_durations(ctx):
    from zerovox_cpp_b200 import capi
    f = np.zeros((3, 528), np.float32)
    with pytest.raises(capi.ZvxError, match="durations"):
        ctx.synth_batch_regulated([f], [np.full(3, -9.0, np.float32)], [np.zeros(528, np.float32)], 1500, False)
