// zvx_internal.h -- host-side declarations shared by the .cu translation units.
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stddef.h>

#include "zvx_common.cuh"
#include "mrf_fused.cuh"

namespace zvx {

// conv_umma.cu -----------------------------------------------------------------------
// Fill p.a_rows / a_stages / b_stages / tmem_cols; returns the dynamic smem bytes.
size_t      conv_umma_plan(ConvParams &p, size_t smem_budget);
cudaError_t conv_umma_init();   // once per device: opt in to > 48 KB dynamic smem
cudaError_t conv_umma_launch(const ConvParams &p, int total_tiles, size_t smem, cudaStream_t st);
// persistent, fully warp-specialised variant (one CTA per SM, double-buffered accumulators)
size_t      conv_umma_pk_plan(ConvParams &p, size_t smem_budget);
cudaError_t conv_umma_pk_launch(const ConvParams &p, int total_tiles, int num_sms, size_t smem, cudaStream_t st);

// conv_ref.cu (validation kernel: plain CUDA cores, same prologue/epilogue arithmetic) ---
cudaError_t conv_ref_launch(const ConvParams &p, int total_tiles, cudaStream_t st);

// mrf_fused.cu (fused MRF residual block, swapped-orientation implicit GEMM) -------------------
cudaError_t mrf_fused_init();   // once per device: opt in to > 48 KB dynamic smem
bool        mrf_fused_supported(int CH, int ncol);
cudaError_t mrf_fused_launch(int CH, const mrf::Params &p, int total_windows, cudaStream_t st);

// aux_kernels.cu ---------------------------------------------------------------------
// per-(utterance, channel) InstanceNorm statistics over the utterance's rows
cudaError_t stats_launch(const float *x, int ld, int ch_off, int C, const int *seg_start, int B, int rate,
                         float *mu, float *rstd, cudaStream_t st);

// finish the statistics started by the conv epilogues (ConvParams::stats_out); channels [C, C + C_tail) are copied
cudaError_t stats_finalize_launch(const double2 *part, int C, const int *tile_start, const int *seg_start, int B, int rate,
                                  const float *tail_mu, const float *tail_rstd, int C_tail, float *mu, float *rstd, cudaStream_t st);

struct AdainDesc {
    const float *fc_w;    // (2C, style_dim) row-major
    const float *fc_b;    // (2C)
    int C;
    int out_off;          // gamma1 at out_off, beta at out_off + C inside one utterance's row
};
constexpr int MAX_ADAIN = 16;
struct AdainTable {
    AdainDesc d[MAX_ADAIN];
    int n;
    int total;            // sum of 2C
    int style_dim;
};
cudaError_t adain_fc_launch(const AdainTable &tab, const float *style, int B, float *out, cudaStream_t st);

// y = ((x-mu)*rstd)*w + b for an [rows][C] fp32 tensor, written to up to two destinations
cudaError_t norm_affine_launch(const float *x, int ldx, int C, const int *seg_start, int B, const float *mu,
                               const float *rstd, const float *w, const float *b, float *dst0, float *dst1, int ld_dst,
                               int dst_ch_off, cudaStream_t st);

// y16 = fp16(lrelu(((x-mu)*rstd)*g + b, slope)) for an [rows][C] slice; stats / affine per utterance
cudaError_t norm_act_f16_launch(const float *x, int ldx, int ch_off, int C, const int *seg_start, int B, int max_len,
                                const float *mu, const float *rstd, const float *g, const float *b, int gb_stride, float slope,
                                __half *y16, __half *raw16 /* optional plain fp16 copy of x */, cudaStream_t st);

cudaError_t sum3_act_f16_launch(const float *a, const float *b, const float *c, float scale, float slope, size_t n, __half *y16,
                                cudaStream_t st);
cudaError_t sum3h_act_f16_launch(const __half *a, const __half *b, const __half *c, float scale, float slope, size_t n, __half *y16,
                                 cudaStream_t st);
cudaError_t length_regulate_launch(const float *feat, const int2 *tab, int n_phonemes, int D, float *out, cudaStream_t st);
cudaError_t cvt_f16_launch(const float *x, int ldx, int ch_off, int C, size_t rows, __half *y16, cudaStream_t st);

// wav = tanh(conv_k(leaky_relu(x, slope)) + b), single output channel
// w_host_kc: host copy of the weights as fp32 [K][C] (constant-bank fast path for C = 32, K = 7), may be null
// x2 / x3 non-null: input = ((x + x2) + x3) * sum_scale (MRF branch average applied by the consumer)
// x16 non-null: the input arrives averaged, activated and rounded to fp16 (stage hand-off of the fused MRF kernel)
// halves != 0: x / x2 / x3 point to fp16 tensors (branch outputs written as fp16 by the fused MRF blocks)
cudaError_t out_conv_launch(const float *x, const float *x2, const float *x3, const __half *x16, int halves, float sum_scale, int C, int K, const __half *w_raw, const float *bias, const float *w_host_kc,
                            float bias_host, float slope, const int *seg_start, const int *tile_start, int B, int rate,
                            int total_tiles, float *wav, int16_t *pcm, cudaStream_t st);

}  // namespace zvx
