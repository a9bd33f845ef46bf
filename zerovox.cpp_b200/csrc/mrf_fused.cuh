// mrf_fused.cuh -- geometry of the fused MRF residual-block kernel (mrf_fused.cu).
//
// What it replaces: HiFiGANResidualBlock, /root/reference/src/hifigan.cpp:74-185 -- for one
// residual block (kernel size k, dilations d_0..d_{P-1}):
//     y = x;  for p:  y = y + conv2_p(lrelu(conv1_p(lrelu(y, .1)) , .1))
// The reference executes 6 ggml_conv_1d (im2col + mul_mat) with fp32 tensors in between;
// this kernel keeps the whole chain of one time window on chip.
//
// Orientation ("swapped" implicit GEMM).  Measured on B200 (tools/ubench/umma_issue.cu):
// tcgen05.mma reads its shared-memory operands at 128 B/clk, so a M=128 x N x K=16 MMA
// costs max(N/2, (4096 + 32 N)/128) cycles: N = 32 / 64 output channels would run at 40 % /
// 67 % of the tensor rate.  Therefore the WEIGHTS are the A operand (M = 128 rows) and the
// time axis is N = 256:
//     D[(s, oc), n'] = sum_{j, ic}  A_j[(s, oc), ic] * X[pos = S n' + j - c, ic]
// with S = 128 / CH output shifts stacked along M (output position S n' + s), taps
// j = s + tap in [0, k + S - 2] and c = (k - 1) / 2.  A_j[(s, oc), :] = W[oc, :, j - s]
// (zero when j - s is not a tap), which is a WINDOW of 128 consecutive rows of the tap-
// reversed weight array -- no stacked copy is materialised.
//
// Positions.  A dilated conv (dilation d) over a window of Wp time steps is a dilation-1
// conv over "positions" when the window is stored phase-major: time tau = m d + r
// (r < d, m < Wp / d) sits at position p = r (Wp / d) + m.  Positions are split into S
// polyphase sub-buffers (q = p mod S, row = p div S) so that a tap shift is a row offset.
// Everything here is shared by the device kernel, the host-side packing code and the CPU
// emulator in tests/cpu/ (which checks exactly this index arithmetic against a direct conv).
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define ZVX_HD __host__ __device__ __forceinline__
#else
#define ZVX_HD inline
#endif

namespace zvx {
namespace mrf {

constexpr int GUARD      = 8;     // zero rows before / after the NCOL data rows of a sub-buffer
constexpr int ROW_BYTES  = 16;    // 8 channels x fp16
constexpr int MAX_LAYERS = 6;
constexpr int MAX_K      = 11;

// rows per (sub-buffer, channel group): 8 + NCOL + 8, +1 so that the 16-byte group segments of
// one warp-wide store land in distinct banks
ZVX_HD constexpr int nrows_of(int ncol) { return ncol + 2 * GUARD + 1; }
ZVX_HD constexpr int lbo_b_of(int ncol) { return nrows_of(ncol) * ROW_BYTES; }
// window length in time steps: the largest multiple of S*d for d in {1,3,5} that fits S*NCOL positions
ZVX_HD constexpr int wp_of(int CH, int ncol) { return ((128 / CH) * ncol / (15 * (128 / CH))) * 15 * (128 / CH); }
// data row used as a dump for elements whose position lies beyond the window (it is only ever read
// by outputs that are themselves beyond the window)
ZVX_HD constexpr int trash_unit(int ncol) { return GUARD + ncol + 4; }

// CH channels, NCOL = GEMM N = columns n' per window (256: one CTA per SM; 128: two CTAs per SM)
template <int CH, int NCOL>
struct Geo {
    static constexpr int S       = 128 / CH;                 // output shifts stacked along M
    static constexpr int NPOS    = S * NCOL;                 // positions per window
    static constexpr int WP      = wp_of(CH, NCOL);
    static constexpr int NROWS   = nrows_of(NCOL);
    static constexpr int LBO_B   = lbo_b_of(NCOL);           // byte distance between channel groups
    static constexpr int GROUPS  = CH / 8;                   // 8-channel groups
    static constexpr int SUB     = GROUPS * LBO_B;           // bytes per polyphase sub-buffer
    static constexpr int BUF     = S * SUB;                  // bytes per activation buffer
    static constexpr int KSTEPS  = CH / 16;                  // MMA K-steps per tap
};

// tap blocks of the reversed, zero-padded weight array of a conv with k taps
ZVX_HD int tap_blocks(int k, int S) { return k + 2 * S - 2; }
// first tap block of the A window for step j (row s of the window uses tap j - s)
ZVX_HD int a_block(int k, int S, int j) { return k + S - 2 - j; }
// bytes of one weight chunk = (layer, 16-channel K-step): two channel groups
ZVX_HD uint32_t chunk_bytes(int k, int S, int CH) { return 2u * (uint32_t)tap_blocks(k, S) * CH * 16u; }

// Channel order inside the kernel.  Accumulator row s*CH + r and K slot r of the activation rows
// hold GLOBAL channel row_to_chan(r): inside every block of 32 channels, row 8 i + g <-> channel
// 4 g + i.  A thread of the tcgen05.ld/st 16x256b fragment layout owns rows g, g+8 (first 16-lane
// half) and g+16, g+24 (second half) of its lane quarter: with this order they are four adjacent
// channels of [time][CH] global memory, so block input and output move as float4 and a warp
// instruction touches whole 128-byte lines.  Weights (rows and K) and biases are packed in row order
// by the host (mrf_fused_host.h); the layers in between never see the permutation.
ZVX_HD int chan_to_row(int c) { return (c & ~31) | ((c & 3) << 3) | ((c >> 2) & 7); }
ZVX_HD int row_to_chan(int r) { return (r & ~31) | ((r & 7) << 2) | ((r >> 3) & 3); }

ZVX_HD int floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

// B operand of step j: sub-buffer and row offset (relative to data row 0)
ZVX_HD void b_step(int k, int S, int j, int &q, int &row_off)
{
    const int u = j - (k - 1) / 2;
    row_off = floor_div(u, S);
    q = u - row_off * S;
}

// time (relative to the window start) of position p in the phase-major layout of dilation d
ZVX_HD int pos_to_tau(int p, int d, int Wp)
{
    const int wd = Wp / d;
    return (p % wd) * d + p / wd;
}
ZVX_HD int tau_to_pos(int tau, int d, int Wp)
{
    const int wd = Wp / d;
    return (tau % d) * wd + tau / d;
}

// Scatter-table entry: where the element (shift s, column n') of a layer's output goes in the
// NEXT layer's input buffer.  unit = 16-byte row index inside the buffer (group 0), tau = its
// time inside the window.
// Packed as: bit 31 valid, bits 17-30 tau, bits 0-16 byte offset of the row (unit * 16).  Entries
// of positions beyond the window are not valid and point at the trash row.
constexpr uint32_t TBL_VALID = 0x80000000u;
ZVX_HD uint32_t tbl_pack(int unit, int tau) { return TBL_VALID | ((uint32_t)tau << 17) | ((uint32_t)unit * 16u); }
ZVX_HD int tbl_unit(uint32_t e) { return (int)((e & 0x1FFFFu) >> 4); }
ZVX_HD uint32_t tbl_byte(uint32_t e) { return e & 0x1FFFFu; }
ZVX_HD int tbl_tau(uint32_t e) { return (int)((e >> 17) & 0x3FFFu); }
ZVX_HD uint32_t tbl_trash(int ncol) { return (uint32_t)trash_unit(ncol) * 16u; }

// unit index of time tau in a buffer laid out for dilation d
ZVX_HD int dest_unit(int tau, int d, int Wp, int S, int groups, int ncol)
{
    const int p = tau_to_pos(tau, d, Wp);
    const int q = p % S, row = p / S;
    return q * groups * nrows_of(ncol) + GUARD + row;
}

struct Layer {
    int             k;          // taps
    int             d;          // dilation (layout of this layer's input and output positions)
    int             accumulate; // 0: conv1 -> H accumulator (fresh); 1: conv2 -> accumulates onto y
    float           out_slope;  // leaky-relu slope applied to (acc + bias) before the fp16 store
    const uint16_t *w;          // packed fp16: [K-step][2 groups][tap block][oc row][8 ic rows] (row order, see chan_to_row)
    const float    *bias;       // [CH] in row order: conv1: its bias; conv2: cumulative sum of conv2 biases so far
    const uint32_t *tbl;        // [S][ncol] scatter into the next layer's buffer (unused for the last)
};

struct Params {
    const float    *y_in;       // [rows][CH] fp32 block input (up-conv output)
    const float    *acc_in;     // running sum over residual blocks or null, indexed like out
    const float    *acc_in2;    // second addend (stage hand-off: the other two blocks' outputs are summed here) or null
    float          *out;        // [rows][CH] fp32 or null
    uint16_t       *out16;      // [rows][CH] fp16(leaky_relu(result, out16_slope)): the next conv's ready-made operand, or null
    float           out16_slope;
    float           scale;      // out = (acc_in + y) * scale when has_scale
    int             has_scale;
    float           in_slope;   // leaky-relu slope of the first conv input (0.1)
    const uint32_t *tbl0;       // [S][ncol] scatter of the block input into layer 0's buffer
    int             nlayers;
    Layer           L[MAX_LAYERS];
    const int      *seg_start;  // [B+1] utterance prefix in frames
    const int      *win_start;  // [B+1] prefix of windows per utterance
    int             B;
    int             ncol;       // 128 or 256 (must match the kernel instantiation)
    int             rate;       // rows per frame at this stage
    int             halo;       // sum of pads of all layers
    int             valid;      // output time steps per window = WP - 2 halo
    int             flags;             // experiment switches (bit 0: scalar prologue also for interior windows)
    int             prefetch;          // 1: L2-prefetch the next window's inputs
    int             resident_ctas;     // persistent grid size (SMs x CTAs per SM); 0: one CTA per window
    int             total_windows;     // filled in by the launcher
    int            *err_flag;
};

}  // namespace mrf
}  // namespace zvx
