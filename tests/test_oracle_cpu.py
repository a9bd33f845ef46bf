"""CPU tests: the oracle restatement against the reference's golden vectors, GGUF layout, host logic."""
import os

import numpy as np
import pytest

import zv_oracle
from conftest import ROOT, golden


def test_instance_norm_matches_reference_known_answer():
    """/root/reference/utils/norm1dexample.json: PyTorch InstanceNorm1d(528, affine) on [1,528,115].
    weight/bias are stored with ~5 digits -> tolerance 2e-4."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "norm1d_example.npz"))
    x = g["x_in"][0].T                      # [T=115, C=528] channels-last
    want = g["x_out"][0].T
    got = zv_oracle.inorm(x) * g["weight"] + g["bias"]
    assert np.abs(got - want).max() < 2e-4


def test_oracle_matches_reference_output_L48(weights, zvx):
    """numpy restatement vs the UNMODIFIED reference (tests/golden/ref_L48.npz, made by make_golden.py).
    Bounds sit just under the reference's own self-noise floor (two ISA builds of the reference
    differ by mel 60.4 dB / wav 61.9 dB; vocoder-only ~65.7 dB)."""
    g = golden(48)
    enc, sty = zvx.synth.make_inputs(48)
    o = zv_oracle.Oracle(weights)
    mel = o.decoder(enc, sty)
    assert zv_oracle.snr_db(g["mel"], mel) >= 55.0
    wav = o.vocoder(g["mel"])
    assert zv_oracle.snr_db(g["wav"], wav) >= 60.0
    assert np.abs(wav - g["wav"]).max() <= 1e-3
    assert wav.shape == (48 * 300,) and np.all(np.abs(wav) < 1.0)


def test_reference_builds_agree_with_each_other():
    """The golden files carry the same reference compiled for two ISAs: their distance is the noise floor
    every parity threshold in this repo is quoted against."""
    for L in (48, 160, 400):
        g = golden(L)
        assert zv_oracle.snr_db(g["wav"], g["wav_v3"]) >= 58.0
        assert zv_oracle.snr_db(g["mel"], g["mel_v3"]) >= 58.0
        assert np.abs(g["wav"] - g["wav_v3"]).max() <= 1e-3


def test_live_reference_reproduces_golden(gguf_path, zvx):
    """If the compiled reference is present (oracle/_ref travels to the GPU box), rerun it."""
    import refrun
    if not refrun.available():
        pytest.skip("oracle/_ref not built")
    g = golden(48)
    enc, sty = zvx.synth.make_inputs(48)
    r = refrun.run(gguf_path, 48, enc, sty)
    assert zv_oracle.snr_db(g["wav"], r["wav"]) >= 58.0
    assert r["wav"].shape == g["wav"].shape


def test_conv_transpose_polyphase_identity():
    """The CUDA path evaluates ConvTranspose1d in polyphase form (zvx_api.cu make_upconv); check the
    index derivation against the reference's zero-stuffing formulation in numpy."""
    rng = np.random.default_rng(3)
    for s, K in ((5, 10), (4, 8), (3, 6)):
        C, OC, T = 16, 8, 23
        x = rng.standard_normal((T, C)).astype(np.float32)
        w = (rng.standard_normal((OC, C, K)) * 0.1).astype(np.float16)
        p, op = s // 2 + s % 2, s % 2
        off = K - 1 - p
        up = np.zeros(((T - 1) * s + 1 + 2 * off + op, C), np.float32)
        up[off:off + (T - 1) * s + 1:s] = x
        want = zv_oracle.conv1d(up, w, None)
        assert want.shape[0] == s * T
        got = np.zeros_like(want)
        xq = zv_oracle.q16(x)
        wf = w.astype(np.float32)
        for phi in range(s):
            k0 = (off - phi) % s
            d0 = (phi + k0 - off) // s
            assert (phi + k0 - off) % s == 0
            for a, k in enumerate(range(k0, K, s)):
                for q in range(T):
                    i = q + d0 + a
                    if 0 <= i < T:
                        got[q * s + phi] += xq[i] @ wf[:, :, k].T
        assert np.abs(got - want).max() < 1e-4


def test_gguf_roundtrip(tmp_path, zvx):
    t = {"a.w": np.arange(24, dtype=np.float16).reshape(2, 3, 4), "a.b": np.arange(5, dtype=np.float32)}
    p = str(tmp_path / "t.gguf")
    zvx.gguf_io.write_gguf(p, {"general.architecture": "x", "x.n": 7}, t)
    kv, r = zvx.gguf_io.read_gguf(p)
    assert kv["x.n"] == 7 and kv["general.architecture"] == "x"
    assert r["a.w"].shape == (2, 3, 4) and np.array_equal(r["a.w"], t["a.w"]) and np.array_equal(r["a.b"], t["a.b"])


def test_model_inventory_follows_zv2gguf_layout(weights):
    """224 hot-path tensors; conv kernels F16 (ggml im2col asserts it), everything else F32."""
    assert len(weights) == 224
    for name, a in weights.items():
        is_conv_w = name.endswith(".w") and a.ndim == 3
        assert a.dtype == (np.float16 if is_conv_w else np.float32), name
    assert weights["_mel_decoder.decode.0.conv1.w"].shape == (1056, 1120, 3)
    assert weights["_mel_decoder.decode.0.norm1.fc.w"].shape == (2240, 528)
    assert weights["_meldec.upsamples.0.1.w"].shape == (256, 512, 10)
    assert weights["_meldec.blocks.11.convs2.2.1.w"].shape == (32, 32, 11)
    assert weights["_meldec.output_conv.1.w"].shape == (1, 32, 7)


def test_lpt_sharding_is_a_balanced_partition(zvx):
    lengths = zvx.synth.batch_lengths(4096, seed=13)
    for world in (1, 2, 4, 8):
        shards = zvx.sharding.shard_utterances(lengths, world)
        allidx = np.sort(np.concatenate(shards))
        assert np.array_equal(allidx, np.arange(4096))
        loads = np.array([lengths[s].sum() for s in shards])
        assert loads.max() - loads.min() <= lengths.max()
    b = zvx.sharding.batches(list(range(100)), lengths, max_utts=64, max_frames=20000)
    assert sum(len(x) for x in b) == 100 and all(len(x) <= 64 for x in b)
    assert all(sum(int(lengths[i]) for i in x) <= 20000 for x in b)
