"""Drop-in test: the reference's own load-and-run sequence (gguf_init_from_file, CPU backend, stages
constructed BEFORE the tensor data is read, decoder->eval, meldec->eval -- zerovox.cpp:28-172,330-334)
with ZeroVOX::StyleTTSDecoder / ZeroVOX::HiFiGAN provided by zerovox.cpp_b200/host/zerovox_b200.{h,cpp}
instead of stylettsdec.cpp / hifigan.cpp.  The binary is built in the container (needs the ggml headers
of the reference tree) and travels to the GPU box."""
import os
import subprocess

import numpy as np
import pytest

import zv_oracle
from conftest import golden, ROOT

pytestmark = pytest.mark.gpu

EXE = os.path.join(ROOT, "zerovox.cpp_b200", "host", "_build", "zvx_dropin")


def test_reference_load_sequence_with_b200_classes(zvx, gguf_path, tmp_path):
    assert os.path.exists(EXE), "zvx_dropin was not built (run __graft_entry__.build() where /root/reference exists)"
    L = 160
    g = golden(L)
    enc, sty = zvx.synth.make_inputs(L)
    enc.tofile(tmp_path / "enc.f32")
    sty.tofile(tmp_path / "sty.f32")
    r = subprocess.run([EXE, gguf_path, str(L), str(tmp_path / "enc.f32"), str(tmp_path / "sty.f32"), str(tmp_path / "out")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    mel = np.fromfile(tmp_path / "out.mel.f32", np.float32).reshape(L, 80)
    wav = np.fromfile(tmp_path / "out.wav.f32", np.float32)
    assert zv_oracle.snr_db(g["mel"], mel) >= 55.0
    assert zv_oracle.snr_db(g["wav"], wav) >= 60.0 and np.abs(wav - g["wav"]).max() <= 1e-3
    # write_wav_file stage (zerovox.cpp:337-391): PCM_16 WAV of the same waveform, libsndfile's conversion
    import wave
    with wave.open(str(tmp_path / "out.wav")) as w:
        assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (1, 2, 24000, L * 300)
        pcm = np.frombuffer(w.readframes(w.getnframes()), "<i2")
    assert np.array_equal(pcm, zv_oracle.pcm16(wav))


def test_missing_tensor_throws_like_checked_get_tensor(zvx, tmp_path):
    """A GGUF without the vocoder tensors: the constructor must fail (reference: utils.cpp:12-15)."""
    assert os.path.exists(EXE)
    t = {k: v for k, v in zvx.synth.make_tensors().items() if not k.startswith("_meldec.blocks.7.")}
    path = str(tmp_path / "broken.gguf")
    zvx.gguf_io.write_gguf(path, zvx.synth.KV, t)
    np.zeros((8, 528), np.float32).tofile(tmp_path / "e.f32")
    np.zeros(528, np.float32).tofile(tmp_path / "s.f32")
    r = subprocess.run([EXE, path, "8", str(tmp_path / "e.f32"), str(tmp_path / "s.f32"), str(tmp_path / "o")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "not found" in r.stderr
