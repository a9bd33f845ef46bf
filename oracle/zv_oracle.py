"""oracle/zv_oracle.py -- TEST INFRASTRUCTURE, not product code.

CPU restatement (numpy, float32 arithmetic) of the reference's mel-decoder + vocoder hot
path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import it;
the product path (libzvx.so) never does.

Parity status: PINNED.  tests/test_oracle.py checks this restatement against
  (a) the reference's only known-answer fixture for this path,
      /root/reference/utils/norm1dexample.json (InstanceNorm1d 528x115), committed as
      tests/golden/norm1d_example.npz, and
  (b) outputs of the UNMODIFIED reference compiled from /root/reference by oracle/Makefile
      (oracle/_ref/zvref_*), committed as tests/golden/ref_L*.npz by
      tests/golden/make_golden.py,
at the reference's own self-noise floor (two builds of the reference for different ISAs
differ by the same amount; see DESIGN.md "Parity floor").

All activations are channels-last [T, C] float32 here (the reference ping-pongs between
[C][L] and [L][C]; the arithmetic is the same).

What defines the numerics:
  * ggml_conv_1d = im2col to F16 + mul_mat with fp32 accumulation
      /root/reference/ggml/src/ggml.c:3769-3786, ggml-cpu/ggml-cpu.c:9890-9961 (fp32->fp16 at :9952),
      :7377-7554 mul_mat, :1463-1503 vec_dot_f16
    -> q(x) = fp16 round of every conv input, fp16 weights, fp32 accumulate.
  * ggml_norm: mean and variance accumulated in double, two-pass, biased, 1/sqrtf(var+eps)
      /root/reference/ggml/src/ggml-cpu/ggml-cpu.c:6880-6929
  * leaky_relu(x,a) = max(x,0) + a*min(x,0)      ggml-cpu.c:1747
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

F32 = np.float32
EPS = 1e-5
UPSAMPLE_SCALES = (5, 5, 4, 3)
RESBLOCK_DILATIONS = (1, 3, 5)
NUM_RESBLOCKS = 3


def q16(x: np.ndarray) -> np.ndarray:
    """fp32 -> fp16 (round to nearest even) -> fp32: what im2col does to every conv input."""
    return x.astype(np.float16).astype(F32)


def lrelu(x: np.ndarray, a: float) -> np.ndarray:
    a = F32(a)
    return (np.maximum(x, F32(0)) + a * np.minimum(x, F32(0))).astype(F32)


def inorm(x: np.ndarray) -> np.ndarray:
    """ggml_norm over the time axis per channel, eps 1e-5 (ggml-cpu.c:6880-6929).
    x [T, C] -> [T, C]."""
    mean = x.astype(np.float64).sum(0) / x.shape[0]
    v = (x - mean.astype(F32)).astype(F32)
    var = (v.astype(np.float64) ** 2).sum(0) / x.shape[0]
    scale = (F32(1.0) / np.sqrt((var.astype(F32) + F32(EPS)).astype(F32))).astype(F32)
    return (v * scale).astype(F32)


def conv1d(x: np.ndarray, w: np.ndarray, b: Optional[np.ndarray], pad: int = 0, dil: int = 1) -> np.ndarray:
    """x [T, IC] fp32, w numpy (OC, IC, K) fp16 (ggml ne [K, IC, OC]) -> [T + 2*pad - dil*(K-1), OC].
    ggml_conv_1d(kernel, data, stride 1, pad, dil) + bias (e.g. hifigan.cpp:132-140)."""
    T, IC = x.shape
    OC, IC2, K = w.shape
    assert IC == IC2
    xq = q16(x)
    xp = np.zeros((T + 2 * pad, IC), F32)
    xp[pad:pad + T] = xq
    OL = T + 2 * pad - dil * (K - 1)
    cols = np.empty((OL, K, IC), F32)
    for k in range(K):
        cols[:, k, :] = xp[k * dil:k * dil + OL]
    wm = w.astype(F32).transpose(2, 1, 0).reshape(K * IC, OC)       # [(k, ic), oc]
    y = cols.reshape(OL, K * IC) @ wm
    if b is not None:
        y = y + b.astype(F32)
    return y.astype(F32)


class Oracle:
    """W: dict name -> numpy array in numpy (reversed-ne) shape, as read_gguf returns."""

    def __init__(self, W: Dict[str, np.ndarray]):
        self.W = W
        self.taps: Dict[str, np.ndarray] = {}
        self.keep_taps = False

    def tap(self, name: str, x: np.ndarray) -> np.ndarray:
        if self.keep_taps:
            self.taps[name] = x
        return x

    def conv(self, x, name, pad=0, dil=1, bias=True):
        return conv1d(x, self.W[name + ".w"], self.W[name + ".b"] if bias else None, pad, dil)

    # ---- StyleTTS decoder ---------------------------------------------------------
    def resblk(self, x, p):
        """ResBlk1d::graph, /root/reference/src/stylettsdec.cpp:69-149."""
        W = self.W
        sc = self.conv(x, p + ".conv1x1", bias=False) if (p + ".conv1x1.w") in W else x
        h = (inorm(x) * W[p + ".norm1.w"]).astype(F32) + W[p + ".norm1.b"]
        h = self.conv(lrelu(h, 0.2), p + ".conv1", pad=1)
        self.tap(p + ".conv1", h)
        h = (inorm(h) * W[p + ".norm2.w"]).astype(F32) + W[p + ".norm2.b"]
        h = self.conv(lrelu(h, 0.2), p + ".conv2", pad=1)
        return self.tap(p, ((h + sc) * F32(1.0 / np.sqrt(2.0))).astype(F32))

    def adain(self, x, s, p):
        """AdaIN1d::graph, stylettsdec.cpp:171-200: (1 + gamma) * IN(x) + beta."""
        W = self.W
        h = (W[p + ".fc.w"].astype(F32) @ s.astype(F32)).astype(F32) + W[p + ".fc.b"]
        C = x.shape[1]
        gamma = (h[:C] + F32(1.0)).astype(F32)
        beta = h[C:]
        return ((inorm(x) * gamma).astype(F32) + beta).astype(F32)

    def adain_resblk(self, x, s, p):
        """AdainResBlk1d::graph, stylettsdec.cpp:242-304 (no upsampling, see :421)."""
        W = self.W
        h = self.conv(lrelu(self.adain(x, s, p + ".norm1"), 0.2), p + ".conv1", pad=1)
        self.tap(p + ".conv1", h)
        h = self.conv(lrelu(self.adain(h, s, p + ".norm2"), 0.2), p + ".conv2", pad=1)
        sc = self.conv(x, p + ".conv1x1", bias=False) if (p + ".conv1x1.w") in W else x
        return self.tap(p, ((h + sc) * F32(1.0 / np.sqrt(2.0))).astype(F32))

    def decoder(self, enc_seq: np.ndarray, style: np.ndarray) -> np.ndarray:
        """StyleTTSDecoder graph, stylettsdec.cpp:371-441. enc_seq [L, 528] -> mel [L, 80]."""
        W = self.W
        x = enc_seq.astype(F32)
        h = self.resblk(self.resblk(x, "_mel_decoder.encode.0"), "_mel_decoder.encode.1")
        a = self.conv(x, "_mel_decoder.asr_res.0")
        a = (inorm(a) * W["_mel_decoder.asr_res.1.w"]).astype(F32) + W["_mel_decoder.asr_res.1.b"]
        self.tap("asr_res", a)
        for i in (0, 1, 2):
            h = self.adain_resblk(np.concatenate([h, a], axis=1), style, f"_mel_decoder.decode.{i}")
        for i in (3, 4):
            h = self.adain_resblk(h, style, f"_mel_decoder.decode.{i}")
        return self.tap("mel", self.conv(h, "_mel_decoder.to_out.0"))

    # ---- HiFi-GAN generator ---------------------------------------------------------
    def conv_transpose(self, c, i, s):
        """conv_transpose1d(), /root/reference/src/hifigan.cpp:22-71: zero-stuff + pad, then a
        stride-1 conv with the pre-flipped kernel (zv2gguf.py:176-178). Intended semantics
        (gaps are zeros; SURVEY.md hazard H1)."""
        name = f"_meldec.upsamples.{i}.1"
        K = self.W[name + ".w"].shape[2]
        p = s // 2 + s % 2
        op = s % 2
        off = K - 1 - p
        T, C = c.shape
        up = np.zeros(((T - 1) * s + 1 + 2 * off + op, C), F32)
        up[off:off + (T - 1) * s + 1:s] = c
        return self.conv(up, name, pad=0)

    def resblock(self, x, idx):
        """HiFiGANResidualBlock(), hifigan.cpp:74-185."""
        y = x
        for d, dil in enumerate(RESBLOCK_DILATIONS):
            n1 = f"_meldec.blocks.{idx}.convs1.{d}.1"
            n2 = f"_meldec.blocks.{idx}.convs2.{d}.1"
            k = self.W[n1 + ".w"].shape[2]
            xt = self.conv(lrelu(y, 0.1), n1, pad=(k - 1) // 2 * dil, dil=dil)
            xt = self.conv(lrelu(xt, 0.1), n2, pad=(k - 1) // 2)
            y = (y + xt).astype(F32)
        return y

    def vocoder(self, mel: np.ndarray) -> np.ndarray:
        """HiFiGAN graph, hifigan.cpp:242-345. mel [L, 80] -> wav [L*300]."""
        W = self.W
        c = ((mel.astype(F32) - W["hifigan.mean"]).astype(F32) / W["hifigan.scale"]).astype(F32)
        c = self.conv(c, "_meldec.input_conv", pad=3)
        self.tap("input_conv", c)
        for i, s in enumerate(UPSAMPLE_SCALES):
            c = self.conv_transpose(lrelu(c, 0.1), i, s)
            self.tap(f"up{i}", c)
            cs = None
            for j in range(NUM_RESBLOCKS):
                y = self.resblock(c, NUM_RESBLOCKS * i + j)
                self.tap(f"block{NUM_RESBLOCKS * i + j}", y)
                cs = y if cs is None else (cs + y).astype(F32)
            c = (cs * F32(1.0 / F32(NUM_RESBLOCKS))).astype(F32)
            self.tap(f"stage{i}", c)
        c = self.conv(lrelu(c, 0.01), "_meldec.output_conv.1", pad=3)
        return np.tanh(c[:, 0]).astype(F32)

    def synth(self, enc_seq, style):
        mel = self.decoder(enc_seq, style)
        return mel, self.vocoder(mel)


def snr_db(ref: np.ndarray, test: np.ndarray) -> float:
    ref = np.asarray(ref, np.float64).ravel()
    d = np.asarray(test, np.float64).ravel() - ref
    den = float((d * d).sum())
    if den == 0.0:
        return float("inf")
    return float(10.0 * np.log10(float((ref * ref).sum()) / den))


# ---------------------------------------------------------------------------------------------
# Output stage.  The reference hands its float waveform to libsndfile
# (/root/reference/src/zerovox.cpp:354-371: SF_FORMAT_WAV | SF_FORMAT_PCM_16, sf_write_float).
# libsndfile is an UNVENDORED, UNPINNED system dependency of the reference (CMakeLists.txt:7-8,
# find_package(SndFile)) and is absent from this image, so its published algorithm is restated:
# src/pcm.c f2s_array -- with the defaults SFC_SET_NORM_FLOAT = SF_TRUE and SFC_SET_CLIPPING =
# SF_FALSE a float sample becomes lrintf(x * 0x7FFF) (round half to even in the default rounding
# mode); src/wav.c writes, for plain PCM, "RIFF" <36 + data bytes> "WAVE" "fmt " <16> {format 1,
# channels, rate, rate * block align, block align, 16 bits} "data" <bytes> and little-endian samples.
# Pinned by tests/test_oracle_cpu.py against Python's own `wave` module and a hand-written vector.
def pcm16(wav):
    return np.rint(np.asarray(wav, np.float32) * np.float32(32767.0)).astype(np.int16)


def wav_file_bytes(pcm, sample_rate):
    import struct
    pcm = np.asarray(pcm, np.int16)
    data = pcm.astype("<i2").tobytes()
    hdr = b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVE" + b"fmt " + struct.pack("<IHHIIHH", 16, 1, 1, sample_rate,
                                                                                         2 * sample_rate, 2, 16)
    return hdr + b"data" + struct.pack("<I", len(data)) + data


# ---------------------------------------------------------------------------------------------
# Length regulator: restatement of the host loop at the end of
# FS2Encoder::eval, /root/reference/src/fs2encoder.cpp:611-654.  PINNED: the unmodified reference program
# (oracle/_ref/zvfull_*: zerovox.cpp + fs2encoder.cpp compiled where they lie) is run on a random-init GGUF that
# carries the FastSpeech2 tensors; its graph outputs (features, log-durations), the expanded hidden_state and the
# returned frame count are committed as tests/golden/regulator_*.npz (generator tests/golden/make_golden.py) and
# tests/test_regulator_cpu.py checks this function against them bit-exactly (default sentence, a random one, and
# one whose expansion hits max_seq_len).
def length_regulate(features, log_dur, max_seq_len):
    """-> (x [max_seq_len][emb] with zero tail, number of valid frames)"""
    import math
    features = np.asarray(features, np.float32)
    x = np.zeros((max_seq_len, features.shape[1]), np.float32)           # :614 memset
    xoff = 0
    for i in range(features.shape[0]):
        dur = np.float32(math.exp(float(np.float32(log_dur[i]))) - 1.0)   # :623 float dur = exp(dur_data[i]) - 1.0
        rounded = int(float(dur) + 0.5)                                   # :624 (int32_t)(dur + 0.5): truncation
        if rounded < 0:
            continue                                                      # :625-626
        for _ in range(rounded):                                          # :631-637
            x[xoff] = features[i]
            xoff += 1
            if xoff >= max_seq_len:
                break
        if xoff >= max_seq_len:                                           # :638-639
            break
    return x, xoff
