#!/usr/bin/env python
"""bench.py -- headline benchmark of the zerovox mel-decoder + HiFi-GAN hot path on B200.

A "step" is one pass of the hot path (StyleTTS decoder + HiFi-GAN generator) over one batch of
64 synthetic utterances of 2-10 s (BASELINE.json configs[1]); with N GPUs every rank processes
its own batch of 64 (utterance sharding, no collective on the data path -> weak scaling).

  value  audio-seconds synthesised per wall-second, inputs already resident in HBM
         (zvx_synth_batch_device), timed with CUDA events on the library's stream.
  e2e    same metric through the host-pointer C-ABI call zvx_synth_batch: pinned host
         inputs -> H2D -> compute -> D2H wav, all inside the timed region.
  roofline  the dominant kernel = conv_umma_kernel on the 72 MRF convolutions
         (82.6 % of the path's FLOPs): algorithmic FLOPs / CUDA-event time of those launches.
  cpu_baseline  the unmodified reference (oracle/_ref) timed on this box's host cores on a
         bounded sample (one 5 s utterance), rank 0, N=1 only.

`--impl reference` times the reference's own CPU implementation (oracle/_ref, all host threads)
on a bounded sample of the same workload and prints the same JSON line with "impl": "reference".
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

METRIC = "audio-sec/sec (RTF^-1), StyleTTS decoder + HiFi-GAN"
UNIT = "audio-s/s"
FLOP_PER_FRAME = 443_835_136          # algorithmic FLOPs per mel frame (true ConvTranspose)
FRAMES_PER_AUDIO_S = 80.0             # 24 kHz / hop 300


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def config_dict(args, world):
    return {
        "workload": "configs[1]: batch of 64 synthetic utterances, L~U{160..800} mel frames (2-10 s), "
                    "features[L,528]+style[528] -> mel[L,80] -> wav[L*300], random-init GGUF (zv2gguf layout, seed 1234)",
        "utterances_per_step_per_gpu": args.batch,
        "length_seed": 11,
        "sharding": f"one pool of {args.batch} x {world} utterances dealt longest-first to the least loaded of {world} rank(s) "
                    f"(sharding.shard_utterances), no collective on the data path",
        "l2": "working set per step (several GB of activations) >> 126 MB L2; no explicit flush",
    }


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile(prefix="clocks_", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 7:
                    continue
                try:
                    sm.append(float(p[0]))
                    mx.append(float(p[1]))
                except ValueError:
                    continue
                for n, v in zip(names, p[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            os.unlink(self.path)
        except Exception:
            pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return j.get("bf16_tflops_sustained", 1400.0), j.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json, sustained)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


def run_reference_sample(L, threads, reps, seed=7):
    """Time the unmodified reference (oracle/_ref) on one utterance of L frames; returns per-rep seconds."""
    from zvxload import zvx
    import refrun
    gguf = zvx.synth.write_model(zvx.synth.default_model_path())
    enc, sty = zvx.synth.make_inputs(L, seed)
    r = refrun.run(gguf, L, enc, sty, threads=threads, reps=reps, want_output=False)
    return r["timing"]["rep_s"], r["binary"]


def reference_arm(args, rank, world):
    if rank != 0:
        return
    from zvxload import zvx
    import refrun
    lengths = zvx.synth.batch_lengths(args.batch, seed=11)
    L = int(sorted(lengths)[len(lengths) // 2])      # the median utterance of the batch
    threads = os.cpu_count() or 1
    cfg = config_dict(args, world)
    cfg["reference_arm_sample"] = (f"ONE CPU process ({threads} host threads) at every N; each step = one utterance of the workload "
                                   f"(the median length of the pool, L={L} frames), not the whole batch: the metric is "
                                   f"normalised (audio-s/s), the batch is the same distribution")
    base = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16xf16->f32",
            "data": "synthetic", "config": cfg}
    if not refrun.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built (needs /root/reference at build time)"}))
        return
    reps, binary = run_reference_sample(L, threads, args.warmup + args.steps)
    timed = reps[args.warmup:]
    total = sum(timed)
    audio_s = L / FRAMES_PER_AUDIO_S
    value = audio_s * len(timed) / total
    sample = (f"each step = the median utterance of the batch (L={L} frames, {audio_s:.2f} s audio) through the "
              f"unmodified reference decoder+vocoder ({binary}), ggml CPU backend")
    base.update({"value": value, "ms_per_step": 1000.0 * total / len(timed), "gpu_launches": 0,
                 "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample},
                 "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(base))


def main():
    args = parse()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and "RANK" not in os.environ:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from zvxload import zvx
    from zerovox_cpp_b200 import capi

    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL writes its banner ("NCCL version ...") and debug lines to stdout by default; stdout carries exactly
        # one JSON line here, so send them to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    gguf = zvx.synth.default_model_path()
    if local_rank == 0:
        zvx.synth.write_model(gguf)
    barrier()
    zvx.synth.write_model(gguf)
    ctx = capi.Context.from_gguf(gguf, device=local_rank)

    # configs[3] design at every N: ONE pool of batch x N utterances, dealt longest-first to the least loaded rank
    # (zerovox.cpp_b200/sharding.py; every rank derives the same assignment from the length list alone)
    pool = zvx.synth.batch_lengths(args.batch * world, seed=11)
    mine = zvx.sharding.shard_utterances(pool, world)[rank]
    lengths = pool[mine]
    B = len(lengths)
    F = int(lengths.sum())
    audio_s = F / FRAMES_PER_AUDIO_S
    Larr = (ctypes.c_int32 * B)(*[int(x) for x in lengths])
    ctx.reserve(F, B)

    g = torch.Generator().manual_seed(1234 + rank)
    h_enc = torch.randn(F, ctx.dim_in, generator=g).pin_memory()
    h_sty = (0.05 * torch.randn(B, ctx.style_dim, generator=g)).pin_memory()
    h_wav = torch.empty(F * ctx.hop, dtype=torch.float32).pin_memory()
    d_enc = h_enc.cuda()
    d_sty = h_sty.cuda()
    d_wav = torch.empty(F * ctx.hop, dtype=torch.float32, device="cuda")
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local_rank))

    def step_device():
        ctx.synth_batch_device(B, d_enc.data_ptr(), d_sty.data_ptr(), Larr, 0, d_wav.data_ptr(), sync=False)

    # ---------------- value: inputs resident in HBM ----------------
    for _ in range(args.warmup):
        step_device()
    ctx.synchronize()
    clocks = ClockSampler(local_rank)
    barrier()
    torch.cuda.synchronize()
    clocks.start()
    launches0 = ctx.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    ctx.synchronize()
    torch.cuda.synchronize()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = ctx.kernel_launches() - launches0
    clk = clocks.stop()
    ms_total = max_over_ranks(ms_total)
    total_audio = sum_over_ranks(audio_s)
    per_rank = [(B, F)]
    if world > 1:
        t = torch.tensor([B, F], dtype=torch.int64, device="cuda")
        gathered = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        per_rank = [(int(x[0]), int(x[1])) for x in gathered]
    value = total_audio * args.steps / (ms_total / 1000.0)

    # ---------------- per-launch CUDA-event timing (separate pass: events between all launches) ----------------
    prof_steps = min(args.steps, 5)
    ctx.profile_begin()
    for _ in range(prof_steps):
        step_device()
    recs = ctx.profile_end()

    # ---------------- e2e: host buffers through the C ABI ----------------
    offs = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
    vp = ctypes.c_void_p
    pe = (vp * B)(*[h_enc.data_ptr() + int(offs[b]) * ctx.dim_in * 4 for b in range(B)])
    ps = (vp * B)(*[h_sty.data_ptr() + b * ctx.style_dim * 4 for b in range(B)])
    pw = (vp * B)(*[h_wav.data_ptr() + int(offs[b]) * ctx.hop * 4 for b in range(B)])
    # headline e2e: the pipelined form of the same call (zvx_synth_batch_submit per step, one zvx_synth_batch_wait at the
    # end): every step still copies its inputs host -> device and its waveform device -> host inside the timed region,
    # but the D2H of step i runs under the H2D + kernels of step i + 1 instead of a host synchronisation per step.
    # Two output buffers alternate, as a consumer that reads batch i while batch i + 1 is produced would have them.
    h_wav2 = torch.empty(F * ctx.hop, dtype=torch.float32).pin_memory()
    pw2 = (vp * B)(*[h_wav2.data_ptr() + int(offs[b]) * ctx.hop * 4 for b in range(B)])
    for i in range(max(1, args.warmup)):
        ctx.synth_batch_submit_ptrs(B, pe, ps, Larr, wav_ptrs=pw2 if i & 1 else pw)
    ctx.synth_batch_wait()
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(args.steps):
        ctx.synth_batch_submit_ptrs(B, pe, ps, Larr, wav_ptrs=pw2 if i & 1 else pw)
    ctx.synth_batch_wait()
    t_e2e = time.perf_counter() - t0
    t_e2e = max_over_ranks(t_e2e)
    e2e_value = total_audio * args.steps / t_e2e
    checksum = float(h_wav[:: 997].double().abs().sum())

    # the synchronous call (returns with the waveform on the host, one host synchronisation per step)
    for _ in range(max(1, args.warmup)):
        ctx.synth_batch_ptrs(B, pe, ps, Larr, None, pw)
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.synth_batch_ptrs(B, pe, ps, Larr, None, pw)
    t_sync = max_over_ranks(time.perf_counter() - t0)
    e2e_sync = {"value": total_audio * args.steps / t_sync, "unit": UNIT, "call": "zvx_synth_batch, one host sync per step"}

    # the reference's caller malloc()s its buffers (zerovox.cpp:63-73): the same pipelined loop with PAGEABLE host memory
    # (the driver stages such copies through its own pinned buffers)
    g_enc = h_enc.clone()                  # plain (pageable) host tensors
    g_sty = h_sty.clone()
    g_wav = [torch.empty(F * ctx.hop, dtype=torch.float32) for _ in range(2)]
    qe = (vp * B)(*[g_enc.data_ptr() + int(offs[b]) * ctx.dim_in * 4 for b in range(B)])
    qs = (vp * B)(*[g_sty.data_ptr() + b * ctx.style_dim * 4 for b in range(B)])
    qw = [(vp * B)(*[w.data_ptr() + int(offs[b]) * ctx.hop * 4 for b in range(B)]) for w in g_wav]
    for i in range(2):
        ctx.synth_batch_submit_ptrs(B, qe, qs, Larr, wav_ptrs=qw[i & 1])
    ctx.synth_batch_wait()
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        ctx.synth_batch_submit_ptrs(B, qe, qs, Larr, wav_ptrs=qw[i & 1])
    ctx.synth_batch_wait()
    t_pg = max_over_ranks(time.perf_counter() - t0)
    e2e_pageable = {"value": total_audio * args.steps / t_pg, "unit": UNIT, "call": "zvx_synth_batch_submit / _wait, malloc'd caller buffers"}

    # same, with the reference's write_wav_file conversion (float -> PCM_16, zerovox.cpp:357-371) done by the output
    # conv on the GPU: zvx_synth_batch_pcm16, half the device -> host bytes (extra key, the headline stays `e2e`)
    h_pcm = torch.empty(F * ctx.hop, dtype=torch.int16).pin_memory()
    pp = (vp * B)(*[h_pcm.data_ptr() + int(offs[b]) * ctx.hop * 2 for b in range(B)])
    for _ in range(max(1, args.warmup)):
        ctx.synth_batch_pcm16_ptrs(B, pe, ps, Larr, pp)
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.synth_batch_pcm16_ptrs(B, pe, ps, Larr, pp)
    t_pcm = max_over_ranks(time.perf_counter() - t0)
    e2e_pcm = {"value": total_audio * args.steps / t_pcm, "unit": UNIT,
               "h2d_bytes_per_step": int(F * ctx.dim_in * 4 + B * ctx.style_dim * 4), "d2h_bytes_per_step": int(F * ctx.hop * 2),
               "pcm_checksum": int(h_pcm[:: 997].long().abs().sum())}

    # same workload entered one stage earlier: phoneme-rate features + log-durations through
    # zvx_synth_batch_regulated (7 frames per phoneme, the last one shorter, so that every utterance expands to the
    # same number of frames as above), PCM_16 out.  Extra key.
    import math
    FPP = 7
    Pn = [(int(L) + FPP - 1) // FPP for L in lengths]
    poffs = np.concatenate([[0], np.cumsum(Pn)]).astype(np.int64)
    h_feat = torch.randn(int(poffs[-1]), ctx.dim_in).pin_memory()
    h_ld = torch.empty(int(poffs[-1]), dtype=torch.float32).pin_memory()
    for b in range(B):
        h_ld[int(poffs[b]):int(poffs[b + 1])] = math.log(FPP + 1.0)
        last = int(lengths[b]) - FPP * (Pn[b] - 1)
        h_ld[int(poffs[b + 1]) - 1] = math.log(last + 1.0)
    pf = (vp * B)(*[h_feat.data_ptr() + int(poffs[b]) * ctx.dim_in * 4 for b in range(B)])
    pl = (vp * B)(*[h_ld.data_ptr() + int(poffs[b]) * 4 for b in range(B)])
    Parr = (ctypes.c_int32 * B)(*Pn)
    maxL = int(max(lengths))
    for b in range(B):
        assert capi.regulated_frames(h_ld[int(poffs[b]):int(poffs[b + 1])].numpy(), maxL) == int(lengths[b])
    for _ in range(max(1, args.warmup)):
        ctx.synth_batch_regulated_ptrs(B, pf, pl, Parr, ps, maxL, False, None, pp)
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.synth_batch_regulated_ptrs(B, pf, pl, Parr, ps, maxL, False, None, pp)
    t_reg = max_over_ranks(time.perf_counter() - t0)
    e2e_reg = {"value": total_audio * args.steps / t_reg, "unit": UNIT,
               "h2d_bytes_per_step": int(poffs[-1]) * (ctx.dim_in * 4 + 4 + 8) + B * ctx.style_dim * 4, "d2h_bytes_per_step": int(F * ctx.hop * 2),
               "phonemes_per_step": int(poffs[-1]), "pcm_checksum": int(h_pcm[:: 997].long().abs().sum())}

    # ---------------- parity of this very build against the committed reference output (configs[0], L = 400) ----------------
    parity = None
    gpath = os.path.join(ROOT, "tests", "golden", "ref_L400.npz")
    if rank == 0 and os.path.exists(gpath):
        gold = np.load(gpath)
        p_enc, p_sty = zvx.synth.make_inputs(400)

        def snr_db(ref, x):
            ref = np.asarray(ref, np.float64).ravel()
            err = np.asarray(x, np.float64).ravel() - ref
            return float(10.0 * np.log10(np.sum(ref * ref) / max(np.sum(err * err), 1e-300)))

        p_mels, p_wavs = ctx.synth_batch([p_enc], [p_sty])
        p_voc = ctx.vocode(gold["mel"])
        parity = {"against": "tests/golden/ref_L400.npz = output of the unmodified reference (oracle/_ref/zvref_native) for "
                             "synth.make_inputs(400), generator tests/golden/make_golden.py",
                  "mel_snr_db": snr_db(gold["mel"], p_mels[0]), "wav_snr_db": snr_db(gold["wav"], p_wavs[0]),
                  "wav_max_abs_err": float(np.abs(p_wavs[0] - gold["wav"]).max()),
                  "vocoder_only_snr_db": snr_db(gold["wav"], p_voc),
                  "vocoder_only_max_abs_err": float(np.abs(p_voc - gold["wav"]).max()),
                  "gates": "wav >= 60 dB and <= 1e-3 max-abs (north star), mel >= 55 dB (tests/test_parity_gpu.py)",
                  "reference_self_floor_db": {"mel": snr_db(gold["mel"], gold["mel_v3"]), "wav": snr_db(gold["wav"], gold["wav_v3"])}}

    # ---------------- roofline of the dominant kernel ----------------
    peak_tf, peak_gbs, peak_src = measured_peaks()
    by = {}
    for kind, stage, flops, nbytes, ms in recs:
        key = kind if kind not in ("mrf_conv", "upconv") else f"{kind}{stage}"
        a = by.setdefault(key, [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += flops
        a[2] += nbytes
        a[3] += ms
    mrf_flops = sum(v[1] for k, v in by.items() if k.startswith("mrf_conv"))
    mrf_ms = sum(v[3] for k, v in by.items() if k.startswith("mrf_conv"))
    mrf_n = sum(v[0] for k, v in by.items() if k.startswith("mrf_conv"))
    achieved = mrf_flops / (mrf_ms / 1000.0) / 1e12 if mrf_ms > 0 else 0.0
    kernel_ms = sum(v[3] for v in by.values())
    breakdown = {k: {"launches": v[0] // prof_steps, "ms_per_step": round(v[3] / prof_steps, 4),
                     "tflops": round(v[1] / (v[3] / 1000.0) / 1e12, 2) if v[3] > 0 and v[1] > 0 else None,
                     "gbs": round(v[2] / (v[3] / 1000.0) / 1e9, 1) if v[3] > 0 and v[2] > 0 else None}
                 for k, v in sorted(by.items())}
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath))
    per_stage = {k: round(v[1] / (v[3] / 1000.0) / 1e12, 1) for k, v in sorted(by.items()) if k.startswith("mrf_conv") and v[3] > 0}
    roofline = {"bound": "tensor",
                "kernel": "the 72 MRF convolutions: mrf_fused_kernel (stages 1-3, 13 fused residual-block launches) + "
                          "conv_umma_pk_kernel (stage 0, 18 launches)",
                "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                "traffic": traffic, "peak_source": peak_src,
                "launches_per_step": mrf_n / max(1, prof_steps), "avg_launch_ms": mrf_ms / max(1, mrf_n),
                "algorithmic_gflop_per_step": mrf_flops / max(1, prof_steps) / 1e9,
                "share_of_step_kernel_time": mrf_ms / kernel_ms if kernel_ms else None,
                "tflops_by_stage": per_stage,
                "whole_path_tflops": value * FLOP_PER_FRAME * FRAMES_PER_AUDIO_S / 1e12 / world}

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f16xf16->f32", "data": "synthetic", "config": config_dict(args, world),
           "clocks": clk, "gpu_launches": launches,
           "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(F * ctx.dim_in * 4 + B * ctx.style_dim * 4),
                   "d2h_bytes_per_step": int(F * ctx.hop * 4), "wav_checksum": checksum},
           "e2e_call": "zvx_synth_batch_submit per step + zvx_synth_batch_wait, pinned caller buffers",
           "e2e_sync": e2e_sync, "e2e_pageable": e2e_pageable,
           "e2e_pcm16": e2e_pcm, "e2e_regulated_pcm16": e2e_reg,
           "parity": parity, "roofline": roofline, "kernel_breakdown": breakdown,
           "audio_s_per_step_per_gpu": audio_s,
           "shards": {"utterances_per_rank": [x[0] for x in per_rank], "frames_per_rank": [x[1] for x in per_rank]}}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            import refrun
            if refrun.available():
                threads = os.cpu_count() or 1
                reps, binary = run_reference_sample(400, threads, 3)
                t = sum(reps[1:]) / len(reps[1:])
                out["cpu_baseline"] = {"value": 5.0 / t, "unit": UNIT, "cores": threads, "kind": "reference",
                                       "sample": f"one 5 s utterance (L=400, configs[0]) through the unmodified reference "
                                                 f"({binary}, ggml CPU), 1 warm-up + 2 timed evals"}
            else:
                out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference",
                                       "sample": "oracle/_ref not built"}
        except Exception as e:  # noqa: BLE001
            out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"failed: {e}"}
    if rank == 0:
        print(json.dumps(out))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
