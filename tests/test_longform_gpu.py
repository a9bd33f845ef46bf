"""BASELINE.json configs[2]: long-form synthesis, 60 s utterances (L = 4800 frames), decoder on the whole
sequence (InstanceNorm is global in time), vocoder halo-tiled.  Checked (a) against the compiled
reference run live on the GPU box's host at L = 4800 when oracle/_ref is present, (b) chunked ==
whole-sequence vocoding (the +-20-frame halo covers the 19.5-frame receptive field),
(c) against the numpy oracle on a shorter long-ish case."""
import numpy as np
import pytest

import zv_oracle

pytestmark = pytest.mark.gpu


def test_chunked_vocoding_equals_whole_sequence(ctx, zvx):
    L = 1500                                                   # the reference's shipped max_seq_len
    enc, sty = zvx.synth.make_inputs(L, seed=3)
    mel = ctx.decode(enc, sty)
    whole = ctx.vocode(mel)
    for chunk, halo in ((256, 20), (100, 20), (333, 32), (4000, 20)):
        got = ctx.vocode_chunked(mel, chunk, halo)
        assert got.shape == whole.shape
        # interior chunk edges see real neighbours, true ends keep the per-layer zero padding: same
        # arithmetic per output sample -> identical up to fp32 accumulation order inside tiles
        assert np.abs(got - whole).max() <= 1e-6, (chunk, halo, float(np.abs(got - whole).max()))
    from zerovox_cpp_b200 import capi
    with pytest.raises(capi.ZvxError, match="halo_frames"):
        ctx.vocode_chunked(mel, 256, 8)


def test_chunk_callbacks_stream_in_order_with_final_data(ctx, zvx):
    """Chunks go through the vocoder in groups (1, 2, 4, 8, ... per pass) but are reported one by one, in order,
    each exactly once, and their samples are final when the callback runs."""
    L = 2600
    enc, sty = zvx.synth.make_inputs(L, seed=9)
    mel = ctx.decode(enc, sty)
    whole = ctx.vocode(mel)
    seen = []

    def on_chunk(first, n, wav):
        assert np.array_equal(wav[first:first + n], whole[first:first + n])
        seen.append((first, n))

    got = ctx.vocode_chunked(mel, 200, 20, on_chunk=on_chunk)
    assert np.array_equal(got, whole)
    assert [f for f, _ in seen] == [i * 200 * 300 for i in range(13)]
    assert sum(n for _, n in seen) == L * 300 and seen[-1][1] == 200 * 300


def test_sixty_second_utterance_matches_live_reference(ctx, zvx, gguf_path):
    import refrun
    if not refrun.available():
        pytest.skip("oracle/_ref not present")
    L = 4800                                                   # 60 s at 80 frames/s
    enc, sty = zvx.synth.make_inputs(L, seed=5)
    r = refrun.run(gguf_path, L, enc, sty)                      # unmodified reference, all host threads
    mel = ctx.decode(enc, sty)
    assert zv_oracle.snr_db(r["mel"], mel) >= 55.0
    voc = ctx.vocode_chunked(r["mel"], 256, 20)                # streamed in overlapping chunks
    assert zv_oracle.snr_db(r["wav"], voc) >= 60.0 and np.abs(voc - r["wav"]).max() <= 1e-3
    wav = ctx.vocode(mel)                                      # end to end, whole sequence
    assert zv_oracle.snr_db(r["wav"], wav) >= 60.0 and np.abs(wav - r["wav"]).max() <= 1e-3
