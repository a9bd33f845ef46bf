"""GPU bring-up diagnostics (not a test): conv kernel unit checks + end-to-end parity numbers."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from zvxload import zvx  # noqa: E402
import zv_oracle  # noqa: E402
from zerovox_cpp_b200 import capi  # noqa: E402

F32 = np.float32


def ref_conv(x, w, bias, pad, dil, rows, pro=None):
    outs = []
    o = 0
    for r in rows:
        xi = x[o:o + r]
        outs.append(zv_oracle.conv1d(xi if pro is None else pro(xi), w, bias, pad, dil))
        o += r
    return np.concatenate(outs, 0)


def conv_cases(ctx):
    rng = np.random.default_rng(0)
    cases = [
        # (rows, Cin, Cout, K, dil)
        ([128], 32, 32, 3, 1),
        ([200, 77], 32, 32, 11, 5),
        ([300], 64, 64, 7, 3),
        ([130, 5], 128, 128, 3, 1),
        ([256], 256, 256, 3, 1),
        ([140], 80, 512, 7, 1),
        ([150], 528, 528, 3, 1),
        ([129], 1120, 1056, 3, 1),
        ([100], 528, 64, 1, 1),
        ([100], 528, 80, 1, 1),
    ]
    for rows, cin, cout, k, dil in cases:
        R = sum(rows)
        x = rng.standard_normal((R, cin)).astype(F32)
        w = (rng.standard_normal((cout, cin, k)) / np.sqrt(cin * k)).astype(np.float16)
        b = rng.standard_normal(cout).astype(F32) * 0.1
        pad = (k - 1) // 2 * dil
        want = ref_conv(x, w, b, pad, dil, rows, pro=lambda a: zv_oracle.lrelu(a, 0.1))
        for validation in (True, False):
            t = time.time()
            try:
                got = ctx.test_conv(rows, x, w, bias=b, dilation=dil, pad=pad, pro_mode=2, pro_slope=0.1,
                                    validation=validation)
                err = np.abs(got - want).max()
                print(f"conv rows={rows} cin={cin} cout={cout} k={k} d={dil} "
                      f"{'validation' if validation else 'umma      '} max|err|={err:.3e} "
                      f"rms={want.std():.3f} ({time.time() - t:.2f}s)", flush=True)
            except Exception as e:  # noqa: BLE001
                print(f"conv rows={rows} cin={cin} cout={cout} k={k} d={dil} validation={validation} FAILED: {e}",
                      flush=True)
                raise


def e2e(ctx, L):
    g = np.load(os.path.join(ROOT, "tests", "golden", f"ref_L{L}.npz"))
    enc, sty = zvx.synth.make_inputs(L)
    for validation in (True, False):
        ctx.set_debug_kernels(validation)
        tag = "validation" if validation else "umma      "
        t = time.time()
        mel = ctx.decode(enc, sty)
        print(f"L={L} {tag} mel   snr={zv_oracle.snr_db(g['mel'], mel):6.2f} dB  max|err|={np.abs(mel - g['mel']).max():.3e}"
              f"  (floor {zv_oracle.snr_db(g['mel'], g['mel_v3']):.2f}) {time.time() - t:.2f}s", flush=True)
        t = time.time()
        wav = ctx.vocode(g["mel"])
        print(f"L={L} {tag} voc   snr={zv_oracle.snr_db(g['wav'], wav):6.2f} dB  max|err|={np.abs(wav - g['wav']).max():.3e}"
              f"  {time.time() - t:.2f}s", flush=True)
        wav2 = ctx.vocode(mel)
        print(f"L={L} {tag} e2e   snr={zv_oracle.snr_db(g['wav'], wav2):6.2f} dB  max|err|={np.abs(wav2 - g['wav']).max():.3e}"
              f"  (floor {zv_oracle.snr_db(g['wav'], g['wav_v3']):.2f})", flush=True)
    ctx.set_debug_kernels(False)


def main():
    gguf = zvx.synth.write_model(zvx.synth.default_model_path())
    t = time.time()
    ctx = capi.Context.from_gguf(gguf)
    print(f"context created in {time.time() - t:.2f}s", flush=True)
    what = sys.argv[1:] or ["conv", "e2e"]
    if "conv" in what:
        conv_cases(ctx)
    if "e2e" in what:
        for L in (48, 160):
            e2e(ctx, L)
    print("launches:", ctx.kernel_launches())


if __name__ == "__main__":
    main()
