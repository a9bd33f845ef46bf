#!/usr/bin/env python
"""BASELINE.json configs[3]: utterance-sharded bulk synthesis -- N synthetic utterances (2-10 s) across
the GPUs of one box.  Launch with torchrun (one rank per GPU) or plain python for one GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 \
        --master-port 29540 tools/bulk_synth.py 4096

Every rank computes the same longest-processing-time-first assignment from the length list alone
(zerovox.cpp_b200/sharding.py), synthesises its shard in launch batches of <= 256 utterances and <= 36000 frames through the
host-pointer C ABI -- every batch is SUBMITTED (zvx_synth_batch_submit: H2D + decoder + vocoder + D2H enqueued, the copies of
one batch run under the kernels of its neighbours) and the rank waits once at the end (zvx_synth_batch_wait); the waveforms land
in one pinned buffer that holds the whole shard -- and rank 0 prints one JSON line.  (`--sync`: one synchronous zvx_synth_batch
per launch batch, as in round 1.)
No collective is on the data path; NCCL is used only for the start barrier and the max-over-ranks time."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from zvxload import zvx  # noqa: E402
from zerovox_cpp_b200 import capi  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    sync_mode = "--sync" in sys.argv
    n_utt = int(args[0]) if args else 4096
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    gguf = zvx.synth.default_model_path()
    if local == 0:
        zvx.synth.write_model(gguf)
    if world > 1:
        dist.barrier()
    ctx = capi.Context.from_gguf(gguf, device=local)
    lengths = zvx.synth.batch_lengths(n_utt, seed=13)
    mine = zvx.sharding.shard_utterances(lengths, world)[rank]
    batches = zvx.sharding.batches(mine, lengths, max_utts=256, max_frames=36000)
    # synthetic inputs of the rank's shard (pinned), generated once and sliced per utterance
    maxL = int(lengths.max())
    g = torch.Generator().manual_seed(99 + rank)
    enc_pool = torch.randn(maxL, ctx.dim_in, generator=g).pin_memory().numpy()
    sty = (0.05 * torch.randn(ctx.style_dim, generator=g)).numpy()
    # pinned output buffer for the whole shard: every batch's waveforms have their own place, so nothing has to be consumed
    # before the next batch may be submitted
    my_frames = int(sum(int(lengths[i]) for i in mine))
    wav_all = torch.empty(my_frames * ctx.hop, dtype=torch.float32).pin_memory()
    import ctypes
    vp = ctypes.c_void_p
    keep = []

    def run_batch(idx, frame0, submit):
        B = len(idx)
        Ls = [int(lengths[i]) for i in idx]
        offs = np.concatenate([[0], np.cumsum(Ls)]).astype(np.int64)
        pe = (vp * B)(*[enc_pool.ctypes.data] * B)          # every utterance reads a prefix of the same pinned pool
        ps = (vp * B)(*[sty.ctypes.data] * B)
        pw = (vp * B)(*[wav_all.data_ptr() + (frame0 + int(offs[b])) * ctx.hop * 4 for b in range(B)])
        La = (ctypes.c_int32 * B)(*Ls)
        if submit:
            keep.append((pe, ps, pw, La))                    # the arrays must outlive the call (the library reads them at submit)
            ctx.synth_batch_submit_ptrs(B, pe, ps, La, pw)
        else:
            ctx.synth_batch_ptrs(B, pe, ps, La, None, pw)
        return int(offs[-1])

    ctx.reserve(36000 + maxL, 256)
    run_batch(batches[0], 0, False)                          # warm-up (workspace, lanes)
    run_batch(batches[-1], 0, False)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    frames = 0
    for b in batches:
        frames += run_batch(b, frames, not sync_mode)
    if not sync_mode:
        ctx.synth_batch_wait()
    dt = time.perf_counter() - t0
    check = float(np.abs(wav_all.numpy()[::9973]).sum())
    if world > 1:
        t = torch.tensor([dt, float(frames)], dtype=torch.float64, device="cuda")
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dt, frames = float(tmax[0]), float(t[1])
    if rank == 0:
        print(json.dumps({"config": "configs[3] utterance-sharded bulk synthesis, pinned host buffers in and out through " +
                                    ("zvx_synth_batch (one synchronous call per launch batch)" if sync_mode else "zvx_synth_batch_submit / _wait (all launch batches in flight)"),
                          "utterances": n_utt, "n_gpus": world, "audio_s": frames / 80.0, "seconds": dt,
                          "audio_s_per_s": frames / 80.0 / dt, "batches_rank0": len(batches), "checksum_rank0": check}))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
