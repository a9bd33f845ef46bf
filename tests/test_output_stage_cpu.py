"""CPU tests of the output stage: the restated libsndfile conversion / WAV layout
(oracle/zv_oracle.py) against known answers and Python's own `wave` reader, and the library's host-only
RIFF writer against both.  No GPU involved."""
import struct
import wave

import numpy as np
import pytest

import zv_oracle


def test_pcm16_conversion_known_answers():
    """libsndfile src/pcm.c f2s_array, normalised, no clipping: lrintf(x * 0x7FFF), ties to even."""
    x = np.array([0.0, 1.0, -1.0, 0.5, -0.5, 1.0 / 32767, 1.5 / 32767, 2.5 / 32767, -1.5 / 32767, 0.25, 0.9999847], np.float32)
    want = [0, 32767, -32767, 16384, -16384, 1, 2, 2, -2, 8192, 32766]
    # 0.5 * 32767 = 16383.5 -> ties to even = 16384; 1.5 -> 2; 2.5 -> 2
    got = zv_oracle.pcm16(x)
    assert got.dtype == np.int16
    ref = [int(np.rint(np.float32(v) * np.float32(32767.0))) for v in x]
    assert got.tolist() == ref
    assert got.tolist()[:5] == want[:5] and got.tolist()[6:9] == want[6:9]


def test_wav_layout_is_read_back_by_an_independent_reader(tmp_path):
    rng = np.random.default_rng(1)
    pcm = (rng.standard_normal(4801) * 9000).clip(-32768, 32767).astype(np.int16)
    b = zv_oracle.wav_file_bytes(pcm, 24000)
    assert b[:4] == b"RIFF" and b[8:16] == b"WAVEfmt " and b[36:40] == b"data"
    assert struct.unpack("<I", b[4:8])[0] == len(b) - 8 and struct.unpack("<I", b[40:44])[0] == 2 * pcm.size
    p = tmp_path / "o.wav"
    p.write_bytes(b)
    with wave.open(str(p)) as w:
        assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (1, 2, 24000, pcm.size)
        assert np.array_equal(np.frombuffer(w.readframes(pcm.size), "<i2"), pcm)


@pytest.mark.parametrize("n", [0, 1, 65536, 70001])
def test_library_wav_writer_matches_the_restated_layout(zvx, tmp_path, n):
    """zvx_write_wav_pcm16 (host only) == oracle bytes, incl. empty and multi-chunk files."""
    from zerovox_cpp_b200 import capi
    rng = np.random.default_rng(n)
    pcm = rng.integers(-32768, 32768, n).astype(np.int16)
    p = str(tmp_path / "w.wav")
    capi.write_wav_pcm16(p, pcm, 24000)
    assert open(p, "rb").read() == zv_oracle.wav_file_bytes(pcm, 24000)
    with pytest.raises(capi.ZvxError):
        capi.write_wav_pcm16(str(tmp_path / "no_such_dir" / "w.wav"), pcm, 24000)
