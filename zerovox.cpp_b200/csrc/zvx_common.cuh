// zvx_common.cuh -- shared device-side definitions for the zerovox B200 hot path.
//
// Layout convention used by every kernel in this library: activations are
// channels-last, row = time step, [rows][ld] with the channel index contiguous
// (the reference ping-pongs between [C][L] and [L][C]; external tensors
// are frame-major [L][C], which is what we keep internally).  A batch of utterances is
// packed back to back along the row axis; `seg` tables give each utterance's row range
// so that the per-layer 'same' zero padding is applied at every true sequence edge
// (SURVEY.md H-d).
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace zvx {

// ---------------------------------------------------------------------------------
// Prologue (what is applied to a conv INPUT element before it is rounded to fp16) and
// epilogue (what happens to the fp32 accumulator) descriptors.  Both the tcgen05
// implicit-GEMM kernel and the plain validation kernel use the same device functions,
// so the per-element arithmetic (and its rounding order) is identical.
// ---------------------------------------------------------------------------------
enum ProMode : int {
    PRO_F16   = 0,   // input already fp16 and already activated: raw copy
    PRO_CVT   = 1,   // fp32 -> fp16
    PRO_LRELU = 2,   // leaky_relu(x, slope) -> fp16              (hifigan.cpp:108,281,324)
    PRO_NORM  = 3,   // ((x-mu)*rstd)*g + b -> leaky_relu(slope)  (stylettsdec.cpp:94-104,191-197,253)
    PRO_MEL   = 4,   // (x - mean) / scale                        (hifigan.cpp:242-243)
    PRO_SUM3H = 6,   // PRO_SUM3 with fp16 sources (own kernel instantiation: the fp32 variant's register budget decides its occupancy)
    PRO_SUM3  = 5,   // leaky_relu(((x + x2) + x3) * sum_scale, slope): the MRF branch sum / average of
                     // hifigan.cpp:300-315 applied by the CONSUMER of the three residual-block outputs
};

struct ConvParams {
    // ---- input ----
    const void  *x;           // fp32 (or fp16 when pro_mode == PRO_F16), [rows][ldx]
    const float *x2, *x3;     // PRO_SUM3: the other two branch outputs, same layout as x
    float        sum_scale;   // PRO_SUM3: 1 / num_resblocks
    int          ldx;         // elements per input row
    int          x_ch_off;    // first input channel inside a row
    int          Cin;
    // ---- batch segmentation (input-rate rows) ----
    const int   *seg_start;   // [B+1] prefix of utterance lengths in FRAMES
    const int   *tile_start;  // [B+1] prefix of M-tiles per utterance at the input rate
    int          B;
    int          rate_in;     // input rows per frame
    // ---- taps: input row for output t, tap a is  t + tap_off0 + a*tap_step ----
    int          ntaps;
    int          tap_off0;
    int          tap_step;
    // ---- weights ----
    const __half *w_packed;   // UMMA-ready blocks (see pack_conv_weights in zvx_api.cu)
    const __half *w_raw;      // original (OC, IC, K) fp16, K fastest (ggml ne [K, IC, OC]); validation kernel
    int          w_taps_total;   // K of the raw tensor
    int          w_tap0;         // raw tap index of tap a is w_tap0 + a*w_tap_stride
    int          w_tap_stride;
    int          Cout;
    int          NC;          // output channels per CTA (N of the UMMA), Cout % NC == 0
    // ---- prologue ----
    int          pro_mode;
    float        pro_slope;
    const float *p_mu;        // [B][Cin]  (PRO_NORM) | [Cin] mean  (PRO_MEL)
    const float *p_rstd;      // [B][Cin]  (PRO_NORM) | [Cin] scale (PRO_MEL)
    int          p_stat_stride;   // per-utterance stride of p_mu / p_rstd (0: shared)
    const float *p_g;         // PRO_NORM multiplicative term
    const float *p_b;         // PRO_NORM additive term
    int          p_gb_stride; // 0: shared by all utterances, else per-utterance stride
    // ---- epilogue:  v = acc + bias; v = v + res; v = acc_in + v; v = v*scale ----
    const float *bias;        // [Cout] or null
    const float *res;         // residual, indexed like out32, or null
    int          ldres;
    int          res_ch_off;
    const float *acc_in;      // running sum (MRF branch accumulation), indexed like out32, or null
    int          has_scale;
    float        scale;
    float       *out32;       // fp32 output or null
    int          ldo32;
    int          o32_ch_off;
    __half      *out16;       // fp16(leaky_relu(v, out16_slope)) output or null
    int          ldo16;
    int          o16_ch_off;
    float        out16_slope;
    int          out_mul;     // output row = seg_out_start + t*out_mul + out_add (polyphase up-conv)
    int          out_add;
    // InstanceNorm statistics of the OUTPUT, fused into the epilogue (one-tile kernel): per (tile, output channel) the sum
    // and the sum of squares of the final values over the tile's valid rows, in double; stats_finalize_kernel turns them
    // into mean / rstd per (utterance, channel).  No atomics: fixed summation order, batch-independent results.
    double2     *stats_out;   // [n_tiles][Cout] or null
    // PRO_F16 operand staged by TMA (cp.async.bulk.tensor.3d, one box = a halo tile of 64 channels) instead of per-thread
    // 16-byte cp.async; the kernel's tensor-map argument describes the [rows][ldx] fp16 buffer as (8 ch, rows, C/8 groups)
    // Second operand source folded into the same accumulator (TMA mode only): Cin_b more input channels read from the fp16
    // matrix xb with ONE tap at the output row -- the learned 1x1 shortcut of a decoder block computed by the block's conv2
    // (D = W2 . h + Wsc . x, stylettsdec.cpp:286-301).  The packed weights of an N-chunk are followed by the shortcut's blocks.
    const void  *xb;
    int          ldxb;
    int          Cin_b;       // 0: none
    int          use_tma;
    int          epi8;        // one-tile kernel, MT = 1: four more epilogue warps (a second warp per tensor-memory lane quarter)
    int          pair;        // one-tile kernel as tcgen05 CTA pairs (.cta_group::2), see conv_umma.cu; needs use_tma
    long long    tma_row0;    // first row of the tensor map inside the buffer (rows are addressed relative to it)
    long long    tma_rows;    // rows of the buffer the tensor map covers
    // ---- smem geometry (host computed) ----
    int          mt;          // M-tiles (128 rows each) per CTA: 1 or 2
    int          a_rows;      // rows per A stage (>= 128*mt + (ntaps-1)*tap_step)
    int          a_stages;
    int          b_stages;
    int          tmem_cols;   // power of two >= max(32, NC)
    int          cluster;     // CTAs per thread-block cluster (1, 2 or 4): they share every weight stage (multicast)
    int          n_tiles;     // tiles of the launch (CTAs beyond it pad the grid to a multiple of `cluster`)
    int         *err_flag;    // device int set on pipeline timeout
};

__device__ __forceinline__ float lrelu_f(float x, float a)
{
    // ggml_vec_leaky_relu_f32 (ggml-cpu.c:1747): max(x,0) + a*min(x,0), separately rounded
    return __fadd_rn(x > 0.f ? x : 0.f, __fmul_rn(a, x < 0.f ? x : 0.f));
}

// per-channel prologue parameters for one channel
struct ProCh { float mu, rstd, g, b; };

__device__ __forceinline__ float prologue_apply(int mode, float x, float slope, const ProCh &pc)
{
    switch (mode) {
        case PRO_LRELU: return lrelu_f(x, slope);
        case PRO_NORM: {
            // ggml_norm: v = x - mean; y = v*scale   (ggml-cpu.c:6914-6922)
            // then ggml_mul, ggml_add (stylettsdec.cpp:97-98 / :195-196), then leaky_relu
            float v = __fsub_rn(x, pc.mu);
            v = __fmul_rn(v, pc.rstd);
            v = __fmul_rn(v, pc.g);
            v = __fadd_rn(v, pc.b);
            return lrelu_f(v, slope);
        }
        case PRO_MEL:   return __fdiv_rn(__fsub_rn(x, pc.mu), pc.rstd);
        default:        return x;
    }
}

__device__ __forceinline__ void epilogue_store(const ConvParams &p, float acc, size_t orow, int oc)
{
    float v = acc;
    if (p.bias)   v = __fadd_rn(v, __ldg(p.bias + oc));
    if (p.res)    v = __fadd_rn(v, p.res[orow * (size_t)p.ldres + p.res_ch_off + oc]);
    if (p.acc_in) v = __fadd_rn(p.acc_in[orow * (size_t)p.ldo32 + p.o32_ch_off + oc], v);
    if (p.has_scale) v = __fmul_rn(v, p.scale);
    if (p.out32)  p.out32[orow * (size_t)p.ldo32 + p.o32_ch_off + oc] = v;
    if (p.out16)  p.out16[orow * (size_t)p.ldo16 + p.o16_ch_off + oc] = __float2half_rn(lrelu_f(v, p.out16_slope));
}

// find utterance u with tile_start[u] <= tile < tile_start[u+1]
__device__ __forceinline__ int find_segment(const int *tile_start, int B, int tile)
{
    int lo = 0, hi = B - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (__ldg(tile_start + mid) <= tile) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// Same search done by a whole (converged) warp with identical arguments: every round the 32 lanes
// probe 32 evenly spaced entries and a ballot picks the sub-range, so a table of 1 + B entries costs
// ceil(log32 B) dependent loads instead of log2 B.  (The tables are re-read by every tile / window;
// between two uses the streaming traffic has evicted them from L1, so each dependent load is an L2
// round trip -- measured as the top stall of the fused MRF prologue, profiles/r01_ncu_mrf_fused_ch32_k3_*.)
__device__ __forceinline__ int find_segment_warp(const int *start, int B, int x)
{
    const int lane = threadIdx.x & 31;
    int lo = 0, n = B;                                   // candidates [lo, lo + n); start[lo] <= x holds
    while (n > 1) {
        const int step = (n + 31) >> 5;
        const int u = lo + lane * step;
        const bool le = u < lo + n && __ldg(start + u) <= x;
        const unsigned m = __ballot_sync(0xffffffffu, le) | 1u;
        const int k = 31 - __clz(m);
        const int end = lo + n;
        lo += k * step;
        n = min(step, end - lo);
    }
    return lo;
}

}  // namespace zvx
