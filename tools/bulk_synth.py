#!/usr/bin/env python
"""BASELINE.json configs[3]: utterance-sharded bulk synthesis -- N synthetic utterances (2-10 s) across
the GPUs of one box.  Launch with torchrun (one rank per GPU) or plain python for one GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 \
        --master-port 29540 tools/bulk_synth.py 4096

Every rank computes the same longest-processing-time-first assignment from the length list alone
(zerovox.cpp_b200/sharding.py), synthesises its shard in launch batches of <= 256 utterances and <= 36000 frames through the
host-pointer C ABI (zvx_synth_batch: H2D + decoder + vocoder + D2H), and rank 0 prints one JSON line.
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
    n_utt = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
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
    # pinned output pool, double-buffered so that a consumer could read batch i while batch i+1 is produced
    wav_pool = [torch.empty(40000 * ctx.hop + maxL * ctx.hop, dtype=torch.float32).pin_memory() for _ in range(2)]
    import ctypes
    vp = ctypes.c_void_p
    turn = [0]

    def run_batch(idx):
        B = len(idx)
        pool = wav_pool[turn[0] & 1]
        turn[0] += 1
        Ls = [int(lengths[i]) for i in idx]
        offs = np.concatenate([[0], np.cumsum(Ls)]).astype(np.int64)
        pe = (vp * B)(*[enc_pool.ctypes.data] * B)          # every utterance reads a prefix of the same pinned pool
        ps = (vp * B)(*[sty.ctypes.data] * B)
        pw = (vp * B)(*[pool.data_ptr() + int(offs[b]) * ctx.hop * 4 for b in range(B)])
        ctx.synth_batch_ptrs(B, pe, ps, (ctypes.c_int32 * B)(*Ls), None, pw)
        return None, [pool.numpy()[: Ls[0] * ctx.hop]]

    ctx.reserve(36000 + maxL, 256)
    run_batch(batches[0])                                   # warm-up (workspace, lanes)
    run_batch(batches[-1])
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    frames = 0
    check = 0.0
    for b in batches:
        _, wavs = run_batch(b)
        frames += int(sum(int(lengths[i]) for i in b))
        check += float(np.abs(wavs[0][::997]).sum())
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt, float(frames)], dtype=torch.float64, device="cuda")
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dt, frames = float(tmax[0]), float(t[1])
    if rank == 0:
        print(json.dumps({"config": "configs[3] utterance-sharded bulk synthesis, pinned host buffers in and out through zvx_synth_batch",
                          "utterances": n_utt, "n_gpus": world, "audio_s": frames / 80.0, "seconds": dt,
                          "audio_s_per_s": frames / 80.0 / dt, "batches_rank0": len(batches), "checksum_rank0": check}))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
