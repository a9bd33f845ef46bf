"""GPU parity of the tcgen05 implicit-GEMM convolution (one layer at a time) against the oracle's conv1d
(numpy: fp16-rounded inputs x fp16 weights, fp32 accumulate) through the C ABI (zvx_test_conv)."""
import numpy as np
import pytest

import zv_oracle

pytestmark = pytest.mark.gpu
F32 = np.float32


def _ref(x, w, bias, pad, dil, rows, pro=None):
    outs, o = [], 0
    for r in rows:
        xi = x[o:o + r]
        outs.append(zv_oracle.conv1d(xi if pro is None else pro(xi), w, bias, pad, dil))
        o += r
    return np.concatenate(outs, 0)


# every (Cin, Cout, K, dilation) class on the hot path (SURVEY.md 8a implicit-GEMM table) + ragged batches
CASES = [
    ([128], 32, 32, 3, 1), ([1], 32, 32, 3, 1), ([127, 129, 3], 32, 32, 7, 3), ([200, 77], 32, 32, 11, 5),
    ([300], 64, 64, 7, 3), ([257], 64, 64, 11, 5), ([130, 5], 128, 128, 3, 1), ([140], 128, 128, 11, 3),
    ([256], 256, 256, 3, 1), ([40, 90], 256, 256, 7, 5), ([140], 80, 512, 7, 1), ([150], 528, 528, 3, 1),
    ([90], 528, 1056, 3, 1), ([70, 60], 1056, 1056, 3, 1), ([129], 1120, 1056, 3, 1), ([65], 1120, 528, 3, 1),
    ([100], 528, 64, 1, 1), ([100], 528, 80, 1, 1), ([33], 528, 1056, 1, 1), ([33], 1120, 1056, 1, 1),
]


@pytest.mark.parametrize("rows,cin,cout,k,dil", CASES)
def test_conv_matches_oracle(ctx, rows, cin, cout, k, dil):
    rng = np.random.default_rng(cin * 7 + cout * 3 + k + dil)
    R = sum(rows)
    x = rng.standard_normal((R, cin)).astype(F32)
    w = (rng.standard_normal((cout, cin, k)) / np.sqrt(cin * k)).astype(np.float16)
    b = (rng.standard_normal(cout) * 0.1).astype(F32)
    pad = (k - 1) // 2 * dil
    want = _ref(x, w, b, pad, dil, rows, pro=lambda a: zv_oracle.lrelu(a, 0.1))
    got = ctx.test_conv(rows, x, w, bias=b, dilation=dil, pad=pad, pro_mode=2, pro_slope=0.1)
    # fp32 accumulation-order differences only: a few ulp of the O(1) outputs
    assert np.abs(got - want).max() <= 3e-5
    val = ctx.test_conv(rows, x, w, bias=b, dilation=dil, pad=pad, pro_mode=2, pro_slope=0.1, validation=True)
    assert np.abs(val - want).max() <= 3e-5


def test_prologue_instance_norm_affine_and_residual_epilogue(ctx):
    """PRO_NORM ((x-mu)*rstd)*g+b -> lrelu(0.2) and epilogue (acc+bias+res)*scale (stylettsdec.cpp:94-146)."""
    rng = np.random.default_rng(5)
    rows, cin, cout = [150, 37], 528, 1056
    R = sum(rows)
    x = (rng.standard_normal((R, cin)) * 2 + 0.3).astype(F32)
    w = (rng.standard_normal((cout, cin, 3)) / np.sqrt(cin * 3)).astype(np.float16)
    b = (rng.standard_normal(cout) * 0.1).astype(F32)
    res = rng.standard_normal((R, cout)).astype(F32)
    mu = rng.standard_normal((2, cin)).astype(F32) * 0.2
    rstd = (0.5 + rng.random((2, cin))).astype(F32)
    g = (1 + 0.1 * rng.standard_normal((2, cin))).astype(F32)
    bb = (0.1 * rng.standard_normal((2, cin))).astype(F32)
    outs, o = [], 0
    for u, r in enumerate(rows):
        xi = x[o:o + r]
        h = ((((xi - mu[u]).astype(F32) * rstd[u]).astype(F32) * g[u]).astype(F32) + bb[u]).astype(F32)
        outs.append(zv_oracle.conv1d(zv_oracle.lrelu(h, 0.2), w, b, 1, 1))
        o += r
    scale = F32(1.0 / np.sqrt(2.0))
    want = ((np.concatenate(outs) + res) * scale).astype(F32)
    got = ctx.test_conv(rows, x, w, bias=b, pad=1, pro_mode=3, pro_slope=0.2, mu=mu, rstd=rstd, g=g, b=bb, res=res,
                        scale=float(scale))
    assert np.abs(got - want).max() <= 5e-5


def test_prologue_mel_normalisation(ctx):
    """PRO_MEL (mel-mean)/scale feeding input_conv k7 (hifigan.cpp:242-265)."""
    rng = np.random.default_rng(6)
    rows, cin, cout = [97], 80, 512
    x = (rng.standard_normal((97, cin)) - 4).astype(F32)
    mean = rng.uniform(-6, -2, cin).astype(F32)
    sc = rng.uniform(0.5, 2, cin).astype(F32)
    w = (rng.standard_normal((cout, cin, 7)) / np.sqrt(cin * 7)).astype(np.float16)
    b = (rng.standard_normal(cout) * 0.1).astype(F32)
    want = zv_oracle.conv1d(((x - mean).astype(F32) / sc).astype(F32), w, b, 3, 1)
    got = ctx.test_conv(rows, x, w, bias=b, pad=3, pro_mode=4, mu=mean, rstd=sc)
    assert np.abs(got - want).max() <= 3e-5


def test_fp16_operand_path_and_fp16_activated_output(ctx):
    """conv1 writes fp16(lrelu(.)) for conv2, which consumes it raw (PRO_F16): hifigan.cpp:150-177."""
    rng = np.random.default_rng(8)
    rows, c = [260, 31], 64
    R = sum(rows)
    x = rng.standard_normal((R, c)).astype(F32)
    w1 = (rng.standard_normal((c, c, 7)) / np.sqrt(c * 7)).astype(np.float16)
    w2 = (rng.standard_normal((c, c, 7)) / np.sqrt(c * 7)).astype(np.float16)
    b1 = (rng.standard_normal(c) * 0.1).astype(F32)
    out1, h16 = ctx.test_conv(rows, x, w1, bias=b1, dilation=3, pad=9, pro_mode=2, pro_slope=0.1, out16_slope=0.1, want16=True)
    want_h16 = zv_oracle.lrelu(out1, 0.1).astype(np.float16)
    assert np.array_equal(h16.view(np.uint16), want_h16.view(np.uint16))          # bit-exact rounding
    y = ctx.test_conv(rows, h16, w2, bias=b1, pad=3, pro_mode=0, res=x)
    want = _ref(h16.astype(F32), w2, b1, 3, 1, rows) + x
    assert np.abs(y - want).max() <= 3e-5


def test_utterance_edges_are_zero_padded_independently(ctx):
    """Packing utterances back to back must not leak across sequence ends (SURVEY.md H-d): batch == singles, bit-exact."""
    rng = np.random.default_rng(9)
    rows, c = [131, 64, 200], 128
    R = sum(rows)
    x = rng.standard_normal((R, c)).astype(F32)
    w = (rng.standard_normal((c, c, 11)) / np.sqrt(c * 11)).astype(np.float16)
    got = ctx.test_conv(rows, x, w, dilation=5, pad=25, pro_mode=2, pro_slope=0.1)
    o = 0
    for r in rows:
        single = ctx.test_conv([r], x[o:o + r], w, dilation=5, pad=25, pro_mode=2, pro_slope=0.1)
        assert np.array_equal(single, got[o:o + r])
        o += r


def test_cta_pair_kernel_equals_single_cta_kernel(zvx, weights, monkeypatch):
    """ZVX_CONV_PAIR=1: the one-tile kernel as tcgen05 CTA pairs (.cta_group::2, M = 256, half a weight stage per SM).  Same
    K order per accumulator, so the result must equal the single-CTA kernel bit for bit -- ragged batch, odd tile count
    (the padding CTA of the last pair stores nothing), partial last K-chunk (1056 = 16 x 64 + 32)."""
    from zerovox_cpp_b200 import capi
    rng = np.random.default_rng(11)
    rows, c, k = [300, 129, 77], 1056, 3
    x = rng.standard_normal((sum(rows), c)).astype(np.float16)
    w = (rng.standard_normal((c, c, k)) / np.sqrt(c * k)).astype(np.float16)
    b = (rng.standard_normal(c) * 0.1).astype(F32)
    outs = []
    for pair in ("0", "1"):
        monkeypatch.setenv("ZVX_CONV_PAIR", pair)
        cx = capi.Context(weights, device=0)          # the switch is read at zvx_create
        outs.append(cx.test_conv(rows, x, w, bias=b, pad=1, pro_mode=0))
        cx.close()
    want = _ref(x.astype(F32), w, b, 1, 1, rows)
    assert np.abs(outs[0] - want).max() <= 5e-5
    assert np.array_equal(outs[0], outs[1])
