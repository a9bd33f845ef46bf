"""GPU tests of the output stage: PCM16 produced by the output conv's epilogue
(zvx_synth_batch_pcm16 / zvx_vocode_pcm16) against the restated libsndfile conversion applied to the float
waveform of the same path (bit-exact: integer work), against the reference's golden waveform, and through
the WAV writer."""
import wave

import numpy as np
import pytest

import zv_oracle
from conftest import golden

pytestmark = pytest.mark.gpu


def test_pcm16_is_the_reference_conversion_of_the_float_waveform_bit_exact(ctx, zvx):
    ins = [zvx.synth.make_inputs(L, seed=20 + i) for i, L in enumerate((48, 5, 160, 333))]
    encs, stys = [e for e, _ in ins], [s for _, s in ins]
    _, wavs = ctx.synth_batch(encs, stys)
    pcms = ctx.synth_batch_pcm16(encs, stys)
    for w, p in zip(wavs, pcms):
        assert p.dtype == np.int16 and p.shape == w.shape
        assert np.array_equal(p, zv_oracle.pcm16(w))


def test_pcm16_matches_reference_golden_within_waveform_tolerance(ctx):
    """1e-3 waveform tolerance (north star) = 33 PCM steps; SNR >= 60 dB on the integer signal as well."""
    g = golden(160)
    pcm = ctx.vocode_pcm16(g["mel"])
    ref = zv_oracle.pcm16(g["wav"])
    assert np.abs(pcm.astype(np.int32) - ref.astype(np.int32)).max() <= 33
    assert zv_oracle.snr_db(ref.astype(np.float64), pcm.astype(np.float64)) >= 58.0      # + quantisation noise of both
    assert np.array_equal(pcm, zv_oracle.pcm16(ctx.vocode(g["mel"])))


def test_wav_file_round_trip(ctx, tmp_path):
    from zerovox_cpp_b200 import capi
    g = golden(48)
    pcm = ctx.vocode_pcm16(g["mel"])
    p = str(tmp_path / "utt.wav")
    capi.write_wav_pcm16(p, pcm, 24000)
    with wave.open(p) as w:
        assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (1, 2, 24000, 48 * 300)
        assert np.array_equal(np.frombuffer(w.readframes(w.getnframes()), "<i2"), pcm)


def test_large_batch_pcm16_uses_both_lanes_and_equals_float_path(ctx, zvx):
    """Batch large enough for the sub-batch / lane pipelining of zvx_synth_batch."""
    lens = zvx.synth.batch_lengths(12, seed=5)
    ins = [zvx.synth.make_inputs(int(L), seed=40 + i) for i, L in enumerate(lens)]
    encs, stys = [e for e, _ in ins], [s for _, s in ins]
    _, wavs = ctx.synth_batch(encs, stys, want_mel=False)
    pcms = ctx.synth_batch_pcm16(encs, stys)
    for w, p in zip(wavs, pcms):
        assert np.array_equal(p, zv_oracle.pcm16(w))
