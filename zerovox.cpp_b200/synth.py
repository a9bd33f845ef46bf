"""Random-init zerovox "medium" model in the zv2gguf.py tensor layout + synthetic inputs.

No real checkpoint exists in this environment (the reference's `medium-ldec.gguf` is
git-ignored, /root/reference/.gitignore:33), so every test and benchmark runs on a
deterministic random-init GGUF with exactly the names / shapes / dtypes the reference
constructors look up:

  decoder  /root/reference/src/stylettsdec.cpp:33-66,163-168,220-239,334-340
  vocoder  /root/reference/src/hifigan.cpp:34-39,123-128,162-167,208-218
  dtypes   /root/reference/utils/zv2gguf.py:156-180 (weight-normed convs folded -> F16,
           everything else F32)
  KV keys  /root/reference/src/zerovox.h:17-33

Determinism: every tensor is drawn from its own PCG64 stream seeded with
crc32(name) ^ seed, so the file is bit-identical on every machine with this numpy.
"""
from __future__ import annotations

import os
import zlib
from typing import Dict, Tuple

import numpy as np

from .gguf_io import write_gguf

ARCH = "zerovox-resnet-fs2-styletts"

EMB_DIM = 512
PUNCT_EMB_DIM = 16
DIM = EMB_DIM + PUNCT_EMB_DIM          # 528
BOTTLENECK = 2 * DIM                   # 1056
RESIDUAL_DIM = 64
NUM_MELS = 80
HOP = 300
SAMPLE_RATE = 24000
UPSAMPLE_SCALES = (5, 5, 4, 3)
UPSAMPLE_KERNELS = (10, 10, 8, 6)
RESBLOCK_KERNELS = (3, 7, 11)
RESBLOCK_DILATIONS = (1, 3, 5)
VOC_CHANNELS = 512

KV = {
    "general.architecture": ARCH,
    f"{ARCH}.max_seq_len": 1500,
    f"{ARCH}.emb_dim": EMB_DIM,
    f"{ARCH}.punct_emb_dim": PUNCT_EMB_DIM,
    f"{ARCH}.decoder.n_head": 2,
    f"{ARCH}.decoder.conv_filter_size": 1024,
    f"{ARCH}.decoder.conv_kernel_size.0": 9,
    f"{ARCH}.decoder.conv_kernel_size.1": 1,
    f"{ARCH}.encoder.layer": 4,
    f"{ARCH}.encoder.head": 2,
    f"{ARCH}.encoder.vp_filter_size": 256,
    f"{ARCH}.encoder.vp_kernel_size": 3,
    f"{ARCH}.encoder.ve_n_bins": 256,
    f"{ARCH}.audio.sampling_rate": SAMPLE_RATE,
    f"{ARCH}.audio.num_mels": NUM_MELS,
    f"{ARCH}.audio.hop_size": HOP,
}


def _rng(name: str, seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64((zlib.crc32(name.encode()) ^ seed) & 0xFFFFFFFF))


def make_tensors(seed: int = 1234) -> Dict[str, np.ndarray]:
    """All 224 hot-path tensors, numpy shape (OC, IC, K) for convs (ggml ne = [K, IC, OC])."""
    t: Dict[str, np.ndarray] = {}

    def conv(name, oc, ic, k, gain=1.0, bias=True):
        w = _rng(name + ".w", seed).standard_normal((oc, ic, k), dtype=np.float32) * (gain / np.sqrt(ic * k))
        t[name + ".w"] = w.astype(np.float16)
        if bias:
            t[name + ".b"] = (_rng(name + ".b", seed).standard_normal(oc, dtype=np.float32) * 0.02).astype(np.float32)

    def affine(name, c):
        t[name + ".w"] = (1.0 + 0.1 * _rng(name + ".w", seed).standard_normal(c, dtype=np.float32)).astype(np.float32)
        t[name + ".b"] = (0.02 * _rng(name + ".b", seed).standard_normal(c, dtype=np.float32)).astype(np.float32)

    def fc(name, c):
        t[name + ".fc.w"] = (_rng(name + ".fc.w", seed).standard_normal((2 * c, DIM), dtype=np.float32)
                             / np.sqrt(DIM)).astype(np.float32)
        t[name + ".fc.b"] = (0.02 * _rng(name + ".fc.b", seed).standard_normal(2 * c, dtype=np.float32)).astype(np.float32)

    t["hifigan.mean"] = _rng("hifigan.mean", seed).uniform(-6.0, -2.0, NUM_MELS).astype(np.float32)
    t["hifigan.scale"] = _rng("hifigan.scale", seed).uniform(0.5, 2.0, NUM_MELS).astype(np.float32)

    # StyleTTS decoder
    for i, (cin, cout) in enumerate(((DIM, BOTTLENECK), (BOTTLENECK, BOTTLENECK))):
        p = f"_mel_decoder.encode.{i}"
        conv(p + ".conv1", cin, cin, 3)
        conv(p + ".conv2", cout, cin, 3)
        if cin != cout:
            conv(p + ".conv1x1", cout, cin, 1, bias=False)
        affine(p + ".norm1", cin)
        affine(p + ".norm2", cin)
    conv("_mel_decoder.asr_res.0", RESIDUAL_DIM, DIM, 1)
    affine("_mel_decoder.asr_res.1", RESIDUAL_DIM)
    cat = BOTTLENECK + RESIDUAL_DIM
    for i, (cin, cout) in enumerate(((cat, BOTTLENECK), (cat, BOTTLENECK), (cat, DIM), (DIM, DIM), (DIM, DIM))):
        p = f"_mel_decoder.decode.{i}"
        conv(p + ".conv1", cout, cin, 3)
        conv(p + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(p + ".conv1x1", cout, cin, 1, bias=False)
        fc(p + ".norm1", cin)
        fc(p + ".norm2", cout)
    conv("_mel_decoder.to_out.0", NUM_MELS, DIM, 1)

    # HiFi-GAN v1 generator
    conv("_meldec.input_conv", VOC_CHANNELS, NUM_MELS, 7)
    ch = VOC_CHANNELS
    for i, (s, k) in enumerate(zip(UPSAMPLE_SCALES, UPSAMPLE_KERNELS)):
        # stored already flipped + permuted to conv form (zv2gguf.py:176-178): numpy (C/2, C, K).
        conv(f"_meldec.upsamples.{i}.1", ch // 2, ch, k)
        ch //= 2
        for j, rk in enumerate(RESBLOCK_KERNELS):
            for d in range(len(RESBLOCK_DILATIONS)):
                conv(f"_meldec.blocks.{3 * i + j}.convs1.{d}.1", ch, ch, rk, gain=0.5)
                conv(f"_meldec.blocks.{3 * i + j}.convs2.{d}.1", ch, ch, rk, gain=0.5)
    conv("_meldec.output_conv.1", 1, ch, 7)
    return t


NUM_PHONEMES = 154
NUM_PUNCTS = 6
MAX_N_PHONEMES = 120
FS2_LAYERS = 4
FS2_FILTER = 1024
FS2_KERNELS = (9, 1)
VP_FILTER = 256
VP_KERNEL = 3
VE_N_BINS = 256
MAX_SEQ_LEN = 1500


def sinusoid_table(n_position: int, d_hid: int) -> np.ndarray:
    """Position-encoding table as zv2gguf.py:41-62 computes it (float32 angles, then sin / cos in place)."""
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    hid = np.arange(d_hid)[None, :]
    tab = (pos / np.power(10000.0, 2 * (hid // 2) / d_hid)).astype(np.float32)
    tab[:, 0::2] = np.sin(tab[:, 0::2])
    tab[:, 1::2] = np.cos(tab[:, 1::2])
    return tab


def make_fs2_tensors(seed: int = 1234) -> Dict[str, np.ndarray]:
    """The FastSpeech2 encoder + variance adaptor tensors FS2Encoder's constructor looks up
    (/root/reference/src/fs2encoder.cpp:29-62,152-171,256-261,344-382,504-505), dtypes per
    zv2gguf.py:156-161 (pos_ffn.w_1/w_2 and the predictors' `conv.w` are F16, the rest F32).  Only the caller side of the
    hot path (length regulator) and the reference oracle use them.  The duration predictor's output layer
    is biased so that a random-init model gives durations of a few frames per phoneme."""
    t: Dict[str, np.ndarray] = {}

    def nrm(name, shape, scale, dtype=np.float32, mean=0.0):
        t[name] = (mean + scale * _rng(name, seed).standard_normal(shape, dtype=np.float32)).astype(dtype)

    nrm("_pe._enc.src_word_emb.w", (NUM_PHONEMES + 1, EMB_DIM), 1.0)
    nrm("_pe._enc.punct_embed.w", (NUM_PUNCTS + 1, PUNCT_EMB_DIM), 1.0)
    t["sinusoid_encoding_table"] = sinusoid_table(MAX_SEQ_LEN + 1, DIM)
    for i in range(FS2_LAYERS):
        p = f"_pe._enc.laystk.{i}"
        for w in ("w_qs", "w_ks", "w_vs", "fc"):
            nrm(f"{p}.slf_attn.{w}.w", (DIM, DIM), 1.0 / np.sqrt(DIM))
            nrm(f"{p}.slf_attn.{w}.b", (DIM,), 0.02)
        nrm(f"{p}.slf_attn.layer_norm.w", (DIM,), 0.1, mean=1.0)
        nrm(f"{p}.slf_attn.layer_norm.b", (DIM,), 0.02)
        nrm(f"{p}.pos_ffn.w_1.w", (FS2_FILTER, DIM, FS2_KERNELS[0]), 1.0 / np.sqrt(DIM * FS2_KERNELS[0]), np.float16)
        nrm(f"{p}.pos_ffn.w_1.b", (FS2_FILTER,), 0.02)
        nrm(f"{p}.pos_ffn.w_2.w", (DIM, FS2_FILTER, FS2_KERNELS[1]), 1.0 / np.sqrt(FS2_FILTER * FS2_KERNELS[1]), np.float16)
        nrm(f"{p}.pos_ffn.w_2.b", (DIM,), 0.02)
        nrm(f"{p}.pos_ffn.layer_norm.w", (DIM,), 0.1, mean=1.0)
        nrm(f"{p}.pos_ffn.layer_norm.b", (DIM,), 0.02)
    for vp, out_bias, out_scale in (("duration_predictor", 1.5, 0.6), ("pitch_predictor", 0.5, 0.25), ("engy_pred", 0.5, 0.25)):
        p = f"_pe._var_adapt.{vp}"
        nrm(f"{p}.conv_layer.conv1d_1.conv.w", (VP_FILTER, DIM, VP_KERNEL), 1.0 / np.sqrt(DIM * VP_KERNEL), np.float16)
        nrm(f"{p}.conv_layer.conv1d_1.conv.b", (VP_FILTER,), 0.02)
        nrm(f"{p}.conv_layer.conv1d_2.conv.w", (VP_FILTER, VP_FILTER, VP_KERNEL), 1.0 / np.sqrt(VP_FILTER * VP_KERNEL), np.float16)
        nrm(f"{p}.conv_layer.conv1d_2.conv.b", (VP_FILTER,), 0.02)
        for k in (1, 2):
            nrm(f"{p}.conv_layer.layer_norm_{k}.w", (VP_FILTER,), 0.1, mean=1.0)
            nrm(f"{p}.conv_layer.layer_norm_{k}.b", (VP_FILTER,), 0.02)
        nrm(f"{p}.linear_layer.w", (1, VP_FILTER), out_scale / np.sqrt(VP_FILTER))
        t[f"{p}.linear_layer.b"] = np.full((1,), out_bias, np.float32)
    nrm("_pe._var_adapt.pitch_embedding.w", (VE_N_BINS, DIM), 0.3)
    nrm("_pe._var_adapt.energy_embedding.w", (VE_N_BINS, DIM), 0.3)
    return t


def write_model(path: str, seed: int = 1234, with_fs2: bool = False) -> str:
    """Write the random-init GGUF to `path` unless it already exists; returns path.  `with_fs2` adds the FastSpeech2
    encoder tensors (the hot-path tensors are identical either way: every tensor has its own stream)."""
    if not os.path.exists(path):
        tmp = f"{path}.tmp.{os.getpid()}"
        tensors = make_tensors(seed)
        if with_fs2:
            tensors.update(make_fs2_tensors(seed))
        write_gguf(tmp, KV, tensors)
        os.replace(tmp, path)
    return path


def default_model_path(seed: int = 1234, with_fs2: bool = False) -> str:
    root = os.environ.get("ZVX_CACHE", "/tmp/zvx_cache")
    os.makedirs(root, exist_ok=True)
    return os.path.join(root, f"zerovox-random-{'full-' if with_fs2 else ''}{seed}.gguf")


def make_inputs(L: int, seed: int = 7) -> Tuple[np.ndarray, np.ndarray]:
    """enc_seq [L, 528] ~ N(0,1) frame-major, style [528] ~ 0.05*N(0,1)."""
    rng = np.random.Generator(np.random.PCG64(seed * 1000003 + L))
    enc = rng.standard_normal((L, DIM), dtype=np.float32)
    style = (0.05 * rng.standard_normal(DIM, dtype=np.float32)).astype(np.float32)
    return enc, style


def batch_lengths(n: int, seed: int = 11, lo: int = 160, hi: int = 800) -> np.ndarray:
    """Utterance lengths in mel frames, U{lo..hi} (2-10 s), SURVEY.md 8d config 2/4."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.integers(lo, hi + 1, size=n).astype(np.int32)
