"""CPU tests of the length regulator row: the restatement of fs2encoder.cpp:611-654 in
oracle/zv_oracle.py on hand-computed cases, and the library's host-side frame count (zvx_regulated_frames)
against it on random and edge inputs.  The restatement itself is pinned by outputs of the unmodified reference program
(tests/golden/regulator_*.npz, made by tests/golden/make_golden.py from oracle/_ref/zvfull_native)."""
import math
import os

import numpy as np
import pytest

import zv_oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def regulator_golden(name):
    g = np.load(os.path.join(GOLDEN, f"regulator_{name}.npz"))
    frames, T = int(g["frames"]), int(g["max_seq_len"])
    hidden = np.zeros((T, g["feat"].shape[1]), np.float32)
    hidden[:frames] = g["feat"][g["frame_src"][:frames]]
    return g, hidden, frames, T


@pytest.mark.parametrize("name", ["default", "random", "capped"])
def test_restatement_pinned_by_reference_outputs(zvx, name):
    """FS2Encoder::eval of the unmodified reference: hidden_state and the returned frame count, bit-exact."""
    from zerovox_cpp_b200 import capi
    g, hidden, frames, T = regulator_golden(name)
    x, n = zv_oracle.length_regulate(g["feat"], g["logdur"], T)
    assert n == frames
    assert np.array_equal(x.view(np.uint32), hidden.view(np.uint32))
    assert capi.regulated_frames(g["logdur"], T) == frames
    if name == "capped":
        assert frames == T          # the expansion was cut at max_seq_len (fs2encoder.cpp:636-640)


def test_restated_regulator_hand_cases():
    f = np.arange(12, dtype=np.float32).reshape(4, 3)
    # durations round(exp(d) - 1): ln(3) -> 2, ln(1.4) -> 0.4 -> 0, ln(2.5) -> 1.5 -> (int)2.0 = 2, ln(1) -> 0
    d = np.array([math.log(3.0), math.log(1.4), math.log(2.5), 0.0], np.float32)
    x, n = zv_oracle.length_regulate(f, d, 10)
    assert n == 4
    assert np.array_equal(x[:4], f[[0, 0, 2, 2]]) and not x[4:].any()
    # cap reached in the middle of a phoneme: the rest is dropped (fs2encoder.cpp:634-639)
    x, n = zv_oracle.length_regulate(f, np.full(4, math.log(4.0), np.float32), 7)
    assert n == 7 and np.array_equal(x, f[[0, 0, 0, 1, 1, 1, 2]])
    # very negative log-duration: exp(d) - 1 -> -1, (int)(-0.5) = 0 frames, never negative
    x, n = zv_oracle.length_regulate(f, np.array([-30.0, -1.0, 0.3, -0.2], np.float32), 10)
    assert n == 0 and not x.any()


def test_library_frame_count_matches_restatement(zvx):
    from zerovox_cpp_b200 import capi
    rng = np.random.default_rng(3)
    for P, cap in ((1, 5), (7, 1500), (40, 100), (300, 1500), (64, 3)):
        for _ in range(20):
            d = rng.normal(1.0, 1.0, P).astype(np.float32)
            _, n = zv_oracle.length_regulate(np.zeros((P, 4), np.float32), d, cap)
            assert capi.regulated_frames(d, cap) == n
    # half-way cases: exp(d) - 1 + 0.5 exactly at / next to an integer
    for k in range(1, 40):
        for eps in (-1e-6, 0.0, 1e-6):
            d = np.array([math.log(k + 0.5 + eps)], np.float32)
            _, n = zv_oracle.length_regulate(np.zeros((1, 4), np.float32), d, 1500)
            assert capi.regulated_frames(d, 1500) == n
    assert capi.regulated_frames(np.array([50.0], np.float32), 1500) == 1500       # huge duration: capped, no overflow
    assert capi.regulated_frames(np.array([np.nan, 1.0], np.float32), 1500) == zv_oracle.length_regulate(
        np.zeros((2, 4), np.float32), np.array([-30.0, 1.0], np.float32), 1500)[1]  # NaN contributes nothing
