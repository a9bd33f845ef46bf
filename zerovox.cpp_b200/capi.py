"""ctypes binding of libzvx.so (include/zvx.h) -- the only way Python reaches the CUDA path.

There is no CPU fallback: if the shared library is missing or no sm_100 device is
present, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import gguf_io

_HERE = os.path.dirname(os.path.abspath(__file__))
# ZVX_LIB: developer override used for A/B timing of two builds inside one GPU session (tools/ab_compare.py)
LIB_PATH = os.environ.get("ZVX_LIB") or os.path.join(_HERE, "libzvx.so")

ZVX_F32, ZVX_F16 = 0, 1


class TensorDesc(C.Structure):
    _fields_ = [("name", C.c_char_p), ("dtype", C.c_int32), ("n_dims", C.c_int32), ("ne", C.c_int64 * 4),
                ("data", C.c_void_p)]


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("dim_in", C.c_int32), ("style_dim", C.c_int32), ("residual_dim", C.c_int32),
                ("num_mels", C.c_int32), ("hop_size", C.c_int32), ("kernel_size", C.c_int32),
                ("num_upsamples", C.c_int32), ("upsample_scales", C.c_int32 * 8), ("num_resblocks", C.c_int32),
                ("num_resblock_dilations", C.c_int32), ("resblock_dilations", C.c_int32 * 32),
                ("with_decoder", C.c_int32), ("with_vocoder", C.c_int32)]


class ConvTest(C.Structure):
    _fields_ = [("B", C.c_int32), ("rows", C.POINTER(C.c_int32)),
                ("Cin", C.c_int32), ("Cout", C.c_int32), ("K", C.c_int32), ("dilation", C.c_int32), ("pad", C.c_int32),
                ("pro_mode", C.c_int32), ("pro_slope", C.c_float),
                ("x", C.c_void_p), ("x16", C.c_void_p), ("w", C.c_void_p), ("bias", C.c_void_p),
                ("mu", C.c_void_p), ("rstd", C.c_void_p), ("g", C.c_void_p), ("b", C.c_void_p),
                ("res", C.c_void_p), ("scale", C.c_float), ("out16_slope", C.c_float),
                ("out", C.c_void_p), ("out16", C.c_void_p), ("use_validation_kernel", C.c_int32)]


class LaunchRecord(C.Structure):
    _fields_ = [("kind", C.c_int32), ("stage", C.c_int32), ("flops", C.c_double), ("bytes", C.c_double),
                ("ms", C.c_float)]


KIND_NAMES = ["dec_conv", "voc_input_conv", "upconv", "mrf_conv", "out_conv", "stats", "adain_fc", "norm_affine"]

EXPORTS = [
    "zvx_default_config", "zvx_create", "zvx_destroy", "zvx_last_error", "zvx_decode", "zvx_vocode",
    "zvx_synth_batch", "zvx_synth_batch_device", "zvx_vocode_batch_device", "zvx_stream", "zvx_synchronize",
    "zvx_kernel_launches", "zvx_reserve", "zvx_set_debug_kernels", "zvx_test_conv", "zvx_debug_fetch",
    "zvx_set_debug_stop", "zvx_profile_begin", "zvx_profile_end", "zvx_set_fused_mrf", "zvx_vocode_batch", "zvx_vocode_chunked",
    "zvx_synth_batch_pcm16", "zvx_vocode_pcm16", "zvx_write_wav_pcm16", "zvx_synth_batch_regulated", "zvx_regulated_frames",
    "zvx_synth_batch_submit", "zvx_synth_batch_wait",
]

_lib = None


def load_library() -> C.CDLL:
    """dlopen libzvx.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not built: run __graft_entry__.build() (nvcc, sm_100a). "
                           "There is no CPU fallback for the hot path.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, f32p = C.c_void_p, C.c_int32, C.c_int64, C.POINTER(C.c_float)
    lib.zvx_default_config.argtypes = [C.POINTER(Config)]
    lib.zvx_default_config.restype = None
    lib.zvx_create.argtypes = [C.POINTER(vp), C.POINTER(Config), C.POINTER(TensorDesc), i32]
    lib.zvx_create.restype = i32
    lib.zvx_destroy.argtypes = [vp]
    lib.zvx_destroy.restype = None
    lib.zvx_last_error.argtypes = [vp]
    lib.zvx_last_error.restype = C.c_char_p
    lib.zvx_decode.argtypes = [vp, vp, vp, i32, vp]
    lib.zvx_decode.restype = i32
    lib.zvx_vocode.argtypes = [vp, vp, i32, vp]
    lib.zvx_vocode.restype = i32
    lib.zvx_synth_batch.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(vp), C.POINTER(i32), C.POINTER(vp), C.POINTER(vp)]
    lib.zvx_synth_batch.restype = i32
    lib.zvx_synth_batch_submit.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(vp), C.POINTER(i32), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    lib.zvx_synth_batch_submit.restype = i32
    lib.zvx_synth_batch_wait.argtypes = [vp]
    lib.zvx_synth_batch_wait.restype = i32
    lib.zvx_vocode_batch.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(i32), C.POINTER(vp)]
    lib.zvx_vocode_batch.restype = i32
    lib.zvx_synth_batch_pcm16.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(vp), C.POINTER(i32), C.POINTER(vp), C.POINTER(vp)]
    lib.zvx_synth_batch_pcm16.restype = i32
    lib.zvx_vocode_pcm16.argtypes = [vp, vp, i32, vp]
    lib.zvx_vocode_pcm16.restype = i32
    lib.zvx_write_wav_pcm16.argtypes = [C.c_char_p, vp, i64, i32]
    lib.zvx_write_wav_pcm16.restype = i32
    lib.zvx_synth_batch_regulated.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(vp), C.POINTER(i32), C.POINTER(vp), i32, i32,
                                              C.POINTER(i32), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    lib.zvx_synth_batch_regulated.restype = i32
    lib.zvx_regulated_frames.argtypes = [vp, i32, i32]
    lib.zvx_regulated_frames.restype = i32
    lib.zvx_vocode_chunked.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp]
    lib.zvx_vocode_chunked.restype = i32
    lib.zvx_synth_batch_device.argtypes = [vp, i32, vp, vp, C.POINTER(i32), vp, vp, i32]
    lib.zvx_synth_batch_device.restype = i32
    lib.zvx_vocode_batch_device.argtypes = [vp, i32, vp, C.POINTER(i32), vp, i32]
    lib.zvx_vocode_batch_device.restype = i32
    lib.zvx_stream.argtypes = [vp]
    lib.zvx_stream.restype = vp
    lib.zvx_synchronize.argtypes = [vp]
    lib.zvx_synchronize.restype = i32
    lib.zvx_kernel_launches.argtypes = [vp]
    lib.zvx_kernel_launches.restype = i64
    lib.zvx_reserve.argtypes = [vp, i64, i32]
    lib.zvx_reserve.restype = i32
    lib.zvx_set_debug_kernels.argtypes = [vp, i32]
    lib.zvx_set_debug_kernels.restype = None
    lib.zvx_set_debug_stop.argtypes = [vp, i32]
    lib.zvx_set_debug_stop.restype = None
    lib.zvx_set_fused_mrf.argtypes = [vp, i32]
    lib.zvx_set_fused_mrf.restype = None
    lib.zvx_test_conv.argtypes = [vp, C.POINTER(ConvTest)]
    lib.zvx_test_conv.restype = i32
    lib.zvx_debug_fetch.argtypes = [vp, C.c_char_p, vp, i64]
    lib.zvx_debug_fetch.restype = i32
    lib.zvx_profile_begin.argtypes = [vp]
    lib.zvx_profile_begin.restype = i32
    lib.zvx_profile_end.argtypes = [vp, C.POINTER(LaunchRecord), i64]
    lib.zvx_profile_end.restype = i64
    _lib = lib
    return lib


def regulated_frames(log_dur: np.ndarray, max_seq_len: int) -> int:
    """zvx_regulated_frames: the valid-frame count FS2Encoder::eval returns (host only)."""
    d = np.ascontiguousarray(log_dur, np.float32)
    return int(load_library().zvx_regulated_frames(d.ctypes.data, d.size, int(max_seq_len)))


def write_wav_pcm16(path: str, pcm: np.ndarray, sample_rate: int) -> None:
    """zvx_write_wav_pcm16: mono 16-bit RIFF/WAVE file (host only)."""
    pcm = np.ascontiguousarray(pcm, np.int16)
    if load_library().zvx_write_wav_pcm16(os.fsencode(path), pcm.ctypes.data, pcm.size, int(sample_rate)) != 0:
        raise ZvxError(f"cannot write {path}")


class ZvxError(RuntimeError):
    pass


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Context:
    """One zvx_ctx (= one GPU).  Mirrors the life cycle of the reference's
    StyleTTSDecoder + HiFiGAN pair (/root/reference/src/zerovox.cpp:119-138)."""

    def __init__(self, weights: Dict[str, np.ndarray], device: int = 0, with_decoder: bool = True,
                 with_vocoder: bool = True):
        self.lib = load_library()
        self.cfg = Config()
        self.lib.zvx_default_config(C.byref(self.cfg))
        self.cfg.device = device
        self.cfg.with_decoder = int(with_decoder)
        self.cfg.with_vocoder = int(with_vocoder)
        descs = (TensorDesc * len(weights))()
        keep = []
        for i, (name, arr) in enumerate(weights.items()):
            arr = np.ascontiguousarray(arr)
            keep.append(arr)
            if arr.dtype == np.float32:
                dt = ZVX_F32
            elif arr.dtype == np.float16:
                dt = ZVX_F16
            else:
                raise ValueError(f"{name}: unsupported dtype {arr.dtype}")
            descs[i].name = name.encode()
            descs[i].dtype = dt
            ne = tuple(reversed(arr.shape))
            descs[i].n_dims = len(ne)
            for k in range(4):
                descs[i].ne[k] = ne[k] if k < len(ne) else 1
            descs[i].data = arr.ctypes.data
        h = C.c_void_p()
        rc = self.lib.zvx_create(C.byref(h), C.byref(self.cfg), descs, len(weights))
        if rc != 0:
            raise ZvxError((self.lib.zvx_last_error(None) or b"?").decode())
        self.h = h
        self.dim_in = self.cfg.dim_in
        self.style_dim = self.cfg.style_dim
        self.num_mels = self.cfg.num_mels
        self.hop = self.cfg.hop_size

    @classmethod
    def from_gguf(cls, path: str, device: int = 0, **kw) -> "Context":
        _, tensors = gguf_io.read_gguf(path)
        return cls(tensors, device=device, **kw)

    def close(self):
        if getattr(self, "h", None):
            self.lib.zvx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise ZvxError((self.lib.zvx_last_error(self.h) or b"?").decode())

    # ---- reference-shaped entry points -------------------------------------------
    def decode(self, enc_seq: np.ndarray, style: np.ndarray) -> np.ndarray:
        enc_seq = np.ascontiguousarray(enc_seq, np.float32)
        style = np.ascontiguousarray(style, np.float32)
        L = enc_seq.shape[0]
        mel = np.empty((L, self.num_mels), np.float32)
        self._check(self.lib.zvx_decode(self.h, _ptr(enc_seq), _ptr(style), L, _ptr(mel)))
        return mel

    def vocode(self, mel: np.ndarray) -> np.ndarray:
        mel = np.ascontiguousarray(mel, np.float32)
        L = mel.shape[0]
        wav = np.empty(L * self.hop, np.float32)
        self._check(self.lib.zvx_vocode(self.h, _ptr(mel), L, _ptr(wav)))
        return wav

    def vocode_chunked(self, mel: np.ndarray, chunk_frames: int = 256, halo_frames: int = 20, on_chunk=None) -> np.ndarray:
        """Long-form vocoding in overlapping mel chunks (configs[2]); equals vocode(mel).
        on_chunk(first_sample, n_samples, wav) is called as soon as those samples are final."""
        mel = np.ascontiguousarray(mel, np.float32)
        L = mel.shape[0]
        wav = np.empty(L * self.hop, np.float32)
        cb = None
        if on_chunk is not None:
            cb = C.CFUNCTYPE(None, C.c_void_p, C.c_int64, C.c_int64)(lambda user, first, n: on_chunk(int(first), int(n), wav))
        self._check(self.lib.zvx_vocode_chunked(self.h, _ptr(mel), L, chunk_frames, halo_frames, _ptr(wav),
                                                C.cast(cb, C.c_void_p) if cb else None, None))
        return wav

    def vocode_batch(self, mel_list: Sequence[np.ndarray]):
        B = len(mel_list)
        mels = [np.ascontiguousarray(m, np.float32) for m in mel_list]
        Ls = (C.c_int32 * B)(*[m.shape[0] for m in mels])
        wavs = [np.empty(m.shape[0] * self.hop, np.float32) for m in mels]
        vp = C.c_void_p
        pm = (vp * B)(*[m.ctypes.data for m in mels])
        pw = (vp * B)(*[w.ctypes.data for w in wavs])
        self._check(self.lib.zvx_vocode_batch(self.h, B, pm, Ls, pw))
        return wavs

    def synth_batch(self, enc_list: Sequence[np.ndarray], style_list: Sequence[np.ndarray], want_mel: bool = True):
        B = len(enc_list)
        encs = [np.ascontiguousarray(e, np.float32) for e in enc_list]
        stys = [np.ascontiguousarray(s, np.float32) for s in style_list]
        Ls = (C.c_int32 * B)(*[e.shape[0] for e in encs])
        mels = [np.empty((e.shape[0], self.num_mels), np.float32) for e in encs] if want_mel else None
        wavs = [np.empty(e.shape[0] * self.hop, np.float32) for e in encs]
        vp = C.c_void_p
        pe = (vp * B)(*[e.ctypes.data for e in encs])
        ps = (vp * B)(*[s.ctypes.data for s in stys])
        pm = (vp * B)(*[m.ctypes.data for m in mels]) if want_mel else None
        pw = (vp * B)(*[w.ctypes.data for w in wavs])
        self._check(self.lib.zvx_synth_batch(self.h, B, pe, ps, Ls, pm, pw))
        return mels, wavs

    def synth_batch_pcm16(self, enc_list: Sequence[np.ndarray], style_list: Sequence[np.ndarray]):
        """Like synth_batch, but the waveforms come back as int16 PCM (zvx_synth_batch_pcm16)."""
        B = len(enc_list)
        encs = [np.ascontiguousarray(e, np.float32) for e in enc_list]
        stys = [np.ascontiguousarray(s, np.float32) for s in style_list]
        Ls = (C.c_int32 * B)(*[e.shape[0] for e in encs])
        pcms = [np.empty(e.shape[0] * self.hop, np.int16) for e in encs]
        vp = C.c_void_p
        pe = (vp * B)(*[e.ctypes.data for e in encs])
        ps = (vp * B)(*[s.ctypes.data for s in stys])
        pp = (vp * B)(*[w.ctypes.data for w in pcms])
        self._check(self.lib.zvx_synth_batch_pcm16(self.h, B, pe, ps, Ls, None, pp))
        return pcms

    def synth_batch_pcm16_ptrs(self, B: int, enc_ptrs, style_ptrs, lengths, pcm_ptrs):
        self._check(self.lib.zvx_synth_batch_pcm16(self.h, B, enc_ptrs, style_ptrs, lengths, None, pcm_ptrs))

    def vocode_pcm16(self, mel: np.ndarray) -> np.ndarray:
        mel = np.ascontiguousarray(mel, np.float32)
        pcm = np.empty(mel.shape[0] * self.hop, np.int16)
        self._check(self.lib.zvx_vocode_pcm16(self.h, mel.ctypes.data, mel.shape[0], pcm.ctypes.data))
        return pcm

    def synth_batch_regulated(self, feats: Sequence[np.ndarray], log_durs: Sequence[np.ndarray], style_list: Sequence[np.ndarray],
                              max_seq_len: int, pad_to_max: bool, pcm16: bool = False, want_mel: bool = False):
        """zvx_synth_batch_regulated: phoneme-rate features + log-durations -> (valid frame counts, waveforms[, mels])."""
        B = len(feats)
        fs = [np.ascontiguousarray(f, np.float32) for f in feats]
        ds = [np.ascontiguousarray(d, np.float32) for d in log_durs]
        stys = [np.ascontiguousarray(s, np.float32) for s in style_list]
        Ps = (C.c_int32 * B)(*[f.shape[0] for f in fs])
        frames = [max_seq_len if pad_to_max else regulated_frames(d, max_seq_len) for d in ds]
        outs = [np.empty(n * self.hop, np.int16 if pcm16 else np.float32) for n in frames]
        vp = C.c_void_p
        pf = (vp * B)(*[f.ctypes.data for f in fs])
        pd = (vp * B)(*[d.ctypes.data for d in ds])
        ps = (vp * B)(*[s.ctypes.data for s in stys])
        po = (vp * B)(*[o.ctypes.data for o in outs])
        valid = (C.c_int32 * B)()
        mels = [np.empty((n, self.num_mels), np.float32) for n in frames] if want_mel else None
        pm = (vp * B)(*[m.ctypes.data for m in mels]) if want_mel else None
        self._check(self.lib.zvx_synth_batch_regulated(self.h, B, pf, pd, Ps, ps, int(max_seq_len), int(pad_to_max), valid, pm,
                                                       None if pcm16 else po, po if pcm16 else None))
        if want_mel:
            return list(valid), outs, mels
        return list(valid), outs

    def synth_batch_regulated_ptrs(self, B: int, feat_ptrs, logdur_ptrs, P, style_ptrs, max_seq_len: int, pad_to_max: bool, wav_ptrs, pcm_ptrs):
        self._check(self.lib.zvx_synth_batch_regulated(self.h, B, feat_ptrs, logdur_ptrs, P, style_ptrs, int(max_seq_len), int(pad_to_max),
                                                       None, None, wav_ptrs, pcm_ptrs))

    def synth_batch_ptrs(self, B: int, enc_ptrs, style_ptrs, lengths, mel_ptrs, wav_ptrs):
        """Raw-pointer form (host pointers as ints) for bench.py: no numpy allocation in the timed region."""
        self._check(self.lib.zvx_synth_batch(self.h, B, enc_ptrs, style_ptrs, lengths, mel_ptrs, wav_ptrs))

    def synth_batch_submit_ptrs(self, B: int, enc_ptrs, style_ptrs, lengths, wav_ptrs=None, pcm_ptrs=None):
        """zvx_synth_batch_submit (raw host pointers): returns once the batch is enqueued; call synth_batch_wait()."""
        self._check(self.lib.zvx_synth_batch_submit(self.h, B, enc_ptrs, style_ptrs, lengths, None, wav_ptrs, pcm_ptrs))

    def synth_batch_wait(self):
        self._check(self.lib.zvx_synth_batch_wait(self.h))

    def synth_batch_device(self, B: int, d_enc: int, d_style: int, lengths, d_mel: int, d_wav: int, sync: bool = False):
        self._check(self.lib.zvx_synth_batch_device(self.h, B, d_enc, d_style, lengths, d_mel or None, d_wav, int(sync)))

    def synchronize(self):
        self._check(self.lib.zvx_synchronize(self.h))

    def stream(self) -> int:
        return int(self.lib.zvx_stream(self.h) or 0)

    def kernel_launches(self) -> int:
        return int(self.lib.zvx_kernel_launches(self.h))

    def reserve(self, total_frames: int, max_batch: int):
        self._check(self.lib.zvx_reserve(self.h, total_frames, max_batch))

    def profile_begin(self):
        self._check(self.lib.zvx_profile_begin(self.h))

    def profile_end(self, max_recs: int = 1 << 16):
        """-> list of (kind_name, stage, flops, bytes, ms), one per kernel launched since profile_begin."""
        recs = (LaunchRecord * max_recs)()
        n = self.lib.zvx_profile_end(self.h, recs, max_recs)
        if n < 0:
            self._check(1)
        return [(KIND_NAMES[r.kind], r.stage, r.flops, r.bytes, r.ms) for r in recs[:min(n, max_recs)]]

    # ---- test / debug surface -------------------------------------------------------
    def set_debug_kernels(self, on: bool):
        self.lib.zvx_set_debug_kernels(self.h, int(on))

    def set_fused_mrf(self, on: bool):
        """True (default): fused residual-block kernel; False: one launch per convolution."""
        self.lib.zvx_set_fused_mrf(self.h, int(on))

    def set_debug_stop(self, stages: int):
        self.lib.zvx_set_debug_stop(self.h, int(stages))

    def debug_fetch(self, what: str, n: int) -> np.ndarray:
        out = np.empty(n, np.float32)
        self._check(self.lib.zvx_debug_fetch(self.h, what.encode(), _ptr(out), n))
        return out

    def test_conv(self, rows: List[int], x: np.ndarray, w: np.ndarray, bias=None, dilation=1, pad=0, pro_mode=1,
                  pro_slope=0.0, mu=None, rstd=None, g=None, b=None, res=None, scale=0.0, out16_slope=0.0,
                  want16=False, validation=False):
        """x [sum rows, Cin] (fp32, or fp16 for pro_mode 0); w (OC, IC, K) fp16."""
        t = ConvTest()
        B = len(rows)
        rows_a = (C.c_int32 * B)(*rows)
        t.B = B
        t.rows = rows_a
        w = np.ascontiguousarray(w, np.float16)
        OC, IC, K = w.shape
        t.Cin, t.Cout, t.K, t.dilation, t.pad = IC, OC, K, dilation, pad
        t.pro_mode, t.pro_slope = pro_mode, pro_slope
        keep = [w]
        if pro_mode == 0:
            x = np.ascontiguousarray(x, np.float16)
            t.x16 = x.ctypes.data
        else:
            x = np.ascontiguousarray(x, np.float32)
            t.x = x.ctypes.data
        keep.append(x)
        t.w = w.ctypes.data

        def f32(a):
            if a is None:
                return None
            a = np.ascontiguousarray(a, np.float32)
            keep.append(a)
            return a.ctypes.data

        t.bias, t.mu, t.rstd, t.g, t.b, t.res = f32(bias), f32(mu), f32(rstd), f32(g), f32(b), f32(res)
        t.scale, t.out16_slope = scale, out16_slope
        R = int(sum(rows))
        out = np.empty((R, OC), np.float32)
        t.out = out.ctypes.data
        out16 = np.empty((R, OC), np.float16) if want16 else None
        t.out16 = out16.ctypes.data if want16 else None
        t.use_validation_kernel = int(validation)
        self._check(self.lib.zvx_test_conv(self.h, C.byref(t)))
        return (out, out16) if want16 else out
