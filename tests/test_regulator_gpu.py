"""GPU tests of the length regulator row: zvx_synth_batch_regulated (phoneme-rate features
expanded on the GPU) must equal zvx_synth_batch fed the matrix the restated host loop (fs2encoder.cpp:611-654)
produces -- bit-exact, the regulator is index work -- in both modes: reference-default (max_seq_len frames, zero
tail) and valid-frames-only."""
import numpy as np
import pytest

import zv_oracle

pytestmark = pytest.mark.gpu


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
