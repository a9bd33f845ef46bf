#!/usr/bin/env python
"""A/B timing of two (or more) library builds / switch settings INSIDE ONE PROCESS, steps interleaved A B A B ..., so
that clock and power state are shared (separate processes on one box differ by up to 10 % on unchanged kernels).

    python tools/ab_interleaved.py [-n ROUNDS] SIDE SIDE [SIDE ...]
    SIDE = name[:lib=path/to/libzvx.so][:ENV=value ...]      (environment switches are read at zvx_create)

Prints ms per (kind, stage) -- median over the rounds of the per-launch CUDA-event times of one bench step -- and the
ratio of every side to the first."""
import ctypes
import importlib.util
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from zvxload import zvx  # noqa: E402


def load_capi(tag, lib):
    spec = importlib.util.spec_from_file_location(f"zerovox_cpp_b200.capi_{tag}", os.path.join(ROOT, "zerovox.cpp_b200", "capi.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    if lib:
        mod.LIB_PATH = os.path.join(ROOT, lib) if not os.path.isabs(lib) else lib
    return mod


def main():
    args = sys.argv[1:]
    rounds = 5
    if args and args[0] == "-n":
        rounds = int(args[1])
        args = args[2:]
    B = 64
    gguf = zvx.synth.write_model(zvx.synth.default_model_path())
    lengths = zvx.synth.batch_lengths(B, seed=11)
    F = int(lengths.sum())
    Larr = (ctypes.c_int32 * B)(*[int(x) for x in lengths])
    sides = []
    for i, a in enumerate(args):
        parts = a.split(":")
        name, lib, env = parts[0], None, {}
        for kv in parts[1:]:
            k, v = kv.split("=", 1)
            if k == "lib":
                lib = v
            else:
                env[k] = v
        old = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        capi = load_capi(f"s{i}", lib)
        ctx = capi.Context.from_gguf(gguf, device=0)
        ctx.reserve(F, B)
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        sides.append((name, ctx))
    d_enc = torch.randn(F, sides[0][1].dim_in, device="cuda")
    d_sty = 0.05 * torch.randn(B, sides[0][1].style_dim, device="cuda")
    d_wav = torch.empty(F * sides[0][1].hop, device="cuda")

    def step(ctx, profile):
        if profile:
            ctx.profile_begin()
        ctx.synth_batch_device(B, d_enc.data_ptr(), d_sty.data_ptr(), Larr, 0, d_wav.data_ptr(), sync=True)
        return ctx.profile_end() if profile else None

    for _ in range(3):
        for _, ctx in sides:
            step(ctx, False)
    acc = [dict() for _ in sides]
    keys = []
    for rnd in range(rounds):
        order = list(range(len(sides)))
        order = order[rnd % len(sides):] + order[:rnd % len(sides)]      # rotate: no side always runs first / last
        for si in order:
            ctx = sides[si][1]
            per = {}
            for kind, stage, flops, nbytes, ms in step(ctx, True):
                k = f"{kind}:st={stage}"
                per[k] = per.get(k, 0.0) + ms
                if k not in keys:
                    keys.append(k)
            per["total"] = sum(per.values())
            # whole step, unprofiled, host clock around a synchronous call (kernels of forked streams overlap: the sum
            # of per-launch times above counts overlapped time twice)
            import time
            ts = []
            for _ in range(2):
                t0 = time.perf_counter()
                step(ctx, False)
                ts.append((time.perf_counter() - t0) * 1e3)
            per["step (host clock)"] = min(ts)
            for k, v in per.items():
                acc[si].setdefault(k, []).append(v)
    print(f"{'median ms over ' + str(rounds) + ' rounds':26s}" + "".join(f"{n:>12s}" for n, _ in sides) + "".join(f"{n + '/' + sides[0][0]:>14s}" for n, _ in sides[1:]))
    for k in keys + ["total", "step (host clock)"]:
        med = [statistics.median(a.get(k, [0.0])) for a in acc]
        # ratio: median over rounds of the per-round ratio (both sides measured back to back in that round)
        rat = [statistics.median([x / y for x, y in zip(a.get(k, [0.0]), acc[0].get(k, [1.0])) if y]) for a in acc[1:]]
        print(f"{k:26s}" + "".join(f"{m:12.4f}" for m in med) + "".join(f"{r:14.3f}" for r in rat))


if __name__ == "__main__":
    main()
