"""CPU tests of the converter row: zerovox.cpp_b200/convert.py restates utils/zv2gguf.py:100-185.
A PyTorch-style state dict (weight_g / weight_v pairs, ConvTranspose1d kernels in torch's (in, out, K) layout, long
names) is converted and checked against PyTorch itself: the weight-norm fold against torch._weight_norm, the
flip + permute of the up-sampling kernels by running torch's own ConvTranspose1d next to the reference's zero-stuffing
formulation (hifigan.cpp:22-71, restated in oracle/zv_oracle.py) on the converted tensor, and names / shapes / dtypes
against the layout the library's loader looks up (synth.make_tensors = what build_decoder / build_vocoder consume)."""
import numpy as np
import torch

import zv_oracle

CFG = {"model": {"max_seq_len": 1500, "emb_dim": 512, "punct_emb_dim": 16,
                 "decoder": {"n_head": 2, "conv_filter_size": 1024, "conv_kernel_size": [9, 1]},
                 "encoder": {"fs2_layer": 4, "fs2_head": 2, "vp_filter_size": 256, "vp_kernel_size": 3, "ve_n_bins": 256}},
       "audio": {"sampling_rate": 24000, "num_mels": 80, "hop_size": 300}}


def _torch_style_checkpoint(zvx, rng):
    """Invert the layout: from the GGUF-side tensors of the synthetic model make what a PyTorch checkpoint would hold."""
    gg = zvx.synth.make_tensors()
    gg.update(zvx.synth.make_fs2_tensors())
    inv = [(s, l) for l, s in zvx.convert.SHORTNAMES]
    sd, gen = {}, {}
    for name, t in gg.items():
        if name in ("hifigan.mean", "hifigan.scale", "sinusoid_encoding_table"):
            continue
        target = gen if name.startswith("_meldec.") else sd
        key = name[len("_meldec."):] if name.startswith("_meldec.") else name
        wn = name.endswith(".w") and t.dtype == np.float16 and (name.startswith("_meldec.") or name.startswith("_mel_decoder."))
        if wn:
            w = t.astype(np.float32)
            if zvx.convert._UPSAMPLE.match(name):
                w = np.ascontiguousarray(np.transpose(w, (1, 0, 2))[:, :, ::-1])     # back to torch's (in, out, K), unflipped
            g = rng.uniform(0.5, 2.0, (w.shape[0],) + (1,) * (w.ndim - 1)).astype(np.float32)
            v = (w * rng.uniform(0.5, 2.0)).astype(np.float32)                       # any v with the same direction ...
            nrm = np.sqrt((v ** 2).sum(axis=tuple(range(1, v.ndim)), keepdims=True))
            g = (np.sqrt((w ** 2).sum(axis=tuple(range(1, w.ndim)), keepdims=True)) + 0 * g).astype(np.float32)   # ... and g = ||w||
            target[key[:-2] + ".weight_v"] = v
            target[key[:-2] + ".weight_g"] = g
            assert np.all(nrm > 0)
        else:
            long = name
            if not name.startswith("_meldec."):
                # undo the shortening (longest replacements first so that 'weight' / 'bias' come back)
                long = name.replace("_pe.", "_phoneme_encoder.").replace("._enc.", "._encoder.").replace("laystk", "layer_stack")
                long = long.replace("_var_adapt", "_variance_adaptor").replace("engy_pred", "energy_predictor")
                if long.endswith(".w"):
                    long = long[:-2] + ".weight"
                elif long.endswith(".b"):
                    long = long[:-2] + ".bias"
                key = long
            else:
                key = key[:-2] + (".weight" if key.endswith(".w") else ".bias")
            target[key] = t.astype(np.float32)
    sd["_meldec.stale"] = np.zeros(3, np.float32)          # the model's own vocoder entries are dropped (zv2gguf.py:100-103)
    sd["some.scalar"] = np.float32(1.0)                      # 0-dim: skipped
    return gg, sd, gen


def test_weight_norm_fold_equals_torch(zvx):
    rng = np.random.default_rng(0)
    for shape in ((32, 16, 3), (64, 1, 7), (8, 24)):
        v = rng.standard_normal(shape).astype(np.float32)
        g = rng.uniform(0.1, 3.0, (shape[0],) + (1,) * (len(shape) - 1)).astype(np.float32)
        ours = zvx.convert.weight_norm_fold(v, g)
        ref = torch._weight_norm(torch.from_numpy(v), torch.from_numpy(g), 0).numpy()
        assert np.allclose(ours, ref, rtol=2e-6, atol=1e-7)
        assert np.array_equal(ours.astype(np.float16), ref.astype(np.float16)) or np.mean(ours.astype(np.float16) != ref.astype(np.float16)) < 1e-3


def test_converted_model_has_the_layout_the_loader_expects(zvx):
    rng = np.random.default_rng(1)
    gg, sd, gen = _torch_style_checkpoint(zvx, rng)
    kv, out = zvx.convert.convert(sd, gen, {"mean": gg["hifigan.mean"], "scale": gg["hifigan.scale"]}, CFG)
    assert kv == zvx.synth.KV
    assert set(out) == set(gg), (sorted(set(out) ^ set(gg))[:8])
    for name, t in gg.items():
        assert out[name].shape == t.shape and out[name].dtype == t.dtype, name
    # F32 tensors pass through unchanged; folded weight-normed kernels come back to their fp16 values (g = ||w||, v || w)
    for name, t in gg.items():
        if t.dtype == np.float32:
            assert np.array_equal(out[name], t), name
    name = "_meldec.blocks.4.convs1.1.1.w"
    assert np.mean(out[name] != gg[name]) < 0.02      # one fp32 rounding in v / ||v||, then fp16: rare last-bit flips
    assert np.allclose(out[name].astype(np.float32), gg[name].astype(np.float32), rtol=2e-3, atol=1e-6)


def test_upsample_kernel_flip_permute_matches_torch_conv_transpose(zvx):
    """ConvTranspose1d(C -> C/2, K, stride s, padding s//2 + s%2, output_padding s%2) in PyTorch == the reference's
    zero-stuffing + stride-1 conv (hifigan.cpp:44-65, oracle restatement) applied to the CONVERTED kernel."""
    rng = np.random.default_rng(2)
    for s, K, C in ((5, 10, 16), (4, 8, 32), (3, 6, 16)):
        w_t = (rng.standard_normal((C, C // 2, K)) / np.sqrt(C * K)).astype(np.float32)      # torch layout (in, out, K)
        b = (0.1 * rng.standard_normal(C // 2)).astype(np.float32)
        g = np.sqrt((w_t ** 2).sum(axis=(1, 2), keepdims=True)).astype(np.float32)
        sd = {}
        gen = {"upsamples.0.1.weight_v": w_t, "upsamples.0.1.weight_g": g, "upsamples.0.1.bias": b}
        _, out = zvx.convert.convert(sd, gen, {"mean": np.zeros(80, np.float32), "scale": np.ones(80, np.float32)}, CFG)
        w_c = out["_meldec.upsamples.0.1.w"]
        assert w_c.shape == (C // 2, C, K) and w_c.dtype == np.float16
        L = 23
        x = rng.standard_normal((L, C)).astype(np.float32)
        p, op = s // 2 + s % 2, s % 2
        w16 = torch.from_numpy(np.transpose(w_c.astype(np.float64), (1, 0, 2))[:, :, ::-1].copy())   # back to torch layout
        xq = torch.from_numpy(x.astype(np.float16).astype(np.float64).T[None])
        ref = torch.nn.functional.conv_transpose1d(xq, w16, torch.from_numpy(b.astype(np.float64)), stride=s, padding=p,
                                                   output_padding=op)[0].numpy().T
        o = zv_oracle.Oracle({"_meldec.upsamples.0.1.w": w_c, "_meldec.upsamples.0.1.b": b})
        ours = o.conv_transpose(x, 0, s)                                   # the reference's zero-stuffing formulation
        assert ours.shape == (L * s, C // 2)
        assert np.allclose(ours, ref, rtol=1e-5, atol=1e-6)
