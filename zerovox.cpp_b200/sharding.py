"""Utterance sharding across the GPUs of one box.

Utterances are independent forward passes and the 125 MB of weights are replicated, so the
only multi-GPU logic on this path is WHO synthesises WHICH utterance: longest-processing-time
first (sort by length, give the next utterance to the least loaded rank).  No collective is
involved; every rank can compute the same assignment from the length list alone.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np


def shard_utterances(lengths: Sequence[int], world: int) -> List[np.ndarray]:
    """-> per-rank arrays of utterance indices (each sorted ascending); deterministic."""
    lengths = np.asarray(lengths, dtype=np.int64)
    if world < 1:
        raise ValueError("world must be >= 1")
    order = np.argsort(-lengths, kind="stable")
    load = np.zeros(world, dtype=np.int64)
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = int(np.argmin(load))
        out[r].append(int(i))
        load[r] += int(lengths[i])
    return [np.array(sorted(o), dtype=np.int64) for o in out]


def batches(indices: Sequence[int], lengths: Sequence[int], max_utts: int = 64, max_frames: int = 40000) -> List[List[int]]:
    """Split one rank's utterances into launch batches bounded in count and total frames."""
    cur: List[int] = []
    frames = 0
    res: List[List[int]] = []
    for i in indices:
        L = int(lengths[i])
        if cur and (len(cur) >= max_utts or frames + L > max_frames):
            res.append(cur)
            cur, frames = [], 0
        cur.append(int(i))
        frames += L
    if cur:
        res.append(cur)
    return res
