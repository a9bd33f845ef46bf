"""Checkpoint -> GGUF converter for the zerovox weight layout.

Restates /root/reference/utils/zv2gguf.py without its two blockers in this environment (h5py for the vocoder's
`stats.h5`, the author's hard-coded checkpoint paths): the inputs are plain name -> array dictionaries (npz files on the
command line, or a PyTorch checkpoint when torch is importable) and the statistics come as an npz / JSON file.

What the reference converter does, rule by rule (zv2gguf.py line numbers):
  * :100-109  every tensor of the vocoder's generator replaces the model's own `_meldec.*` entries;
  * :117-139  fifteen uint32 hyper-parameters from the model configuration;
  * :141-142  `hifigan.mean` / `hifigan.scale` from the vocoder statistics;
  * :149-151  0-dim tensors are skipped;
  * :22-39    names are shortened by plain substring replacement, in this order;
  * :156-161  `pos_ffn.w_1.w`, `pos_ffn.w_2.w` and every `*conv.w` become F16;
  * :164-180  `weight_g` is dropped, `weight_v` is folded with torch._weight_norm(v, g, 0) = v * g / ||v|| (norm over all
              dims but 0), named `<long name with weight_v -> w>` (NOT shortened) and stored F16; the transposed-conv
              kernels `_meldec.upsamples.N.1.w` are additionally flipped along the tap axis and permuted to
              (out, in, K) so that a plain stride-1 conv over the zero-stuffed signal applies them (hifigan.cpp:22-71);
  * :184-185  the sinusoid position table of max_seq_len + 1 rows.
Everything else is written as it comes (F32).  gguf stores dims fastest-first, i.e. numpy (OC, IC, K) -> ne [K, IC, OC].
"""
from __future__ import annotations

import json
import re
import sys
from typing import Dict, Mapping, Tuple

import numpy as np

from .gguf_io import write_gguf
from .synth import ARCH, sinusoid_table

SHORTNAMES = (("_phoneme_encoder", "_pe"), ("_encoder", "_enc"), ("layer_stack", "laystk"), ("weight", "w"),
              ("_variance_adaptor", "_var_adapt"), ("energy_predictor", "engy_pred"), ("bias", "b"))
_UPSAMPLE = re.compile(r"^_meldec.upsamples.[0-9].1.w$")


def shorten_tensor_name(long_name: str) -> str:
    s = long_name
    for long, short in SHORTNAMES:
        s = s.replace(long, short)
    return s


def weight_norm_fold(v: np.ndarray, g: np.ndarray) -> np.ndarray:
    """torch._weight_norm(v, g, dim=0): w = v * (g / ||v||), the norm taken over every dim except 0, in float32."""
    v = np.asarray(v, np.float32)
    g = np.asarray(g, np.float32)
    axes = tuple(range(1, v.ndim))
    norm = np.sqrt(np.sum(v.astype(np.float32) ** 2, axis=axes, keepdims=True, dtype=np.float32)).astype(np.float32)
    return (v * (g.reshape(norm.shape) / norm)).astype(np.float32)


def kv_from_config(cfg: Mapping) -> Dict[str, object]:
    m, a = cfg["model"], cfg["audio"]
    return {
        "general.architecture": ARCH,
        f"{ARCH}.max_seq_len": m["max_seq_len"],
        f"{ARCH}.emb_dim": m["emb_dim"],
        f"{ARCH}.punct_emb_dim": m["punct_emb_dim"],
        f"{ARCH}.decoder.n_head": m["decoder"]["n_head"],
        f"{ARCH}.encoder.layer": m["encoder"]["fs2_layer"],
        f"{ARCH}.encoder.head": m["encoder"]["fs2_head"],
        f"{ARCH}.encoder.vp_filter_size": m["encoder"]["vp_filter_size"],
        f"{ARCH}.encoder.vp_kernel_size": m["encoder"]["vp_kernel_size"],
        f"{ARCH}.encoder.ve_n_bins": m["encoder"]["ve_n_bins"],
        f"{ARCH}.decoder.conv_filter_size": m["decoder"]["conv_filter_size"],
        f"{ARCH}.decoder.conv_kernel_size.0": m["decoder"]["conv_kernel_size"][0],
        f"{ARCH}.decoder.conv_kernel_size.1": m["decoder"]["conv_kernel_size"][1],
        f"{ARCH}.audio.sampling_rate": a["sampling_rate"],
        f"{ARCH}.audio.num_mels": a["num_mels"],
        f"{ARCH}.audio.hop_size": a["hop_size"],
    }


def convert(state_dict: Mapping[str, np.ndarray], generator: Mapping[str, np.ndarray], stats: Mapping[str, np.ndarray],
            cfg: Mapping) -> Tuple[Dict[str, object], Dict[str, np.ndarray]]:
    """-> (kv, tensors) ready for gguf_io.write_gguf, in the reference converter's order."""
    sd = {k: np.asarray(v) for k, v in state_dict.items() if not k.startswith("_meldec.")}
    for k, v in generator.items():
        sd["_meldec." + k] = np.asarray(v)
    out: Dict[str, np.ndarray] = {}
    out["hifigan.mean"] = np.asarray(stats["mean"], np.float32)
    out["hifigan.scale"] = np.asarray(stats["scale"], np.float32)
    for key in sorted(sd):
        t = sd[key]
        if t.ndim == 0:
            continue
        if key.endswith("weight_g"):
            continue
        if key.endswith("weight_v"):
            w = weight_norm_fold(t, sd[key.replace(".weight_v", ".weight_g")])
            name = key.replace("weight_v", "w")
            if _UPSAMPLE.match(name):
                w = np.ascontiguousarray(np.transpose(w[:, :, ::-1], (1, 0, 2)))     # flip taps, then (in, out, K) -> (out, in, K)
            out[name] = w.astype(np.float16)
            continue
        name = shorten_tensor_name(key)
        if name.endswith("pos_ffn.w_1.w") or name.endswith("pos_ffn.w_2.w") or name.endswith("conv.w"):
            t = t.astype(np.float16)
        elif t.dtype not in (np.float32, np.float16):
            t = t.astype(np.float32)
        out[name] = np.ascontiguousarray(t)
    m = cfg["model"]
    out["sinusoid_encoding_table"] = sinusoid_table(m["max_seq_len"] + 1, m["emb_dim"] + m["punct_emb_dim"])
    return kv_from_config(cfg), out


def _load_arrays(path: str) -> Dict[str, np.ndarray]:
    if path.endswith(".npz"):
        with np.load(path) as z:
            return {k: z[k] for k in z.files}
    if path.endswith(".json"):
        with open(path) as f:
            return {k: np.asarray(v, np.float32) for k, v in json.load(f).items()}
    import torch                       # .ckpt / .pkl checkpoints, only where torch is available
    obj = torch.load(path, map_location="cpu", weights_only=False)
    for k in ("state_dict", "model"):
        if isinstance(obj, dict) and k in obj:
            obj = obj[k]
    if isinstance(obj, dict) and "generator" in obj:
        obj = obj["generator"]
    return {k: v.detach().cpu().numpy() for k, v in obj.items()}


def main(argv=None) -> int:
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) != 5:
        print("usage: python -m zerovox_cpp_b200.convert model_state.{npz,ckpt} vocoder_generator.{npz,pkl} stats.{npz,json} "
              "modelcfg.{yaml,json} out.gguf", file=sys.stderr)
        return 2
    sd, gen, st = _load_arrays(argv[0]), _load_arrays(argv[1]), _load_arrays(argv[2])
    if argv[3].endswith(".json"):
        cfg = json.load(open(argv[3]))
    else:
        import yaml
        cfg = yaml.safe_load(open(argv[3]))
    kv, tensors = convert(sd, gen, st, cfg)
    write_gguf(argv[4], kv, tensors)
    print(f"{argv[4]} written: {len(tensors)} tensors")
    return 0


if __name__ == "__main__":
    sys.exit(main())
