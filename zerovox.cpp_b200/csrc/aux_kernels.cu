// aux_kernels.cu -- the non-GEMM kernels of the hot path: InstanceNorm statistics, AdaIN
// style projection, affine-norm apply (asr_res branch) and the single-channel output conv.
// All are HBM/L2-bound; they use coalesced channel-contiguous accesses and shuffle / smem
// reductions.
#include "zvx_common.cuh"
#include "zvx_internal.h"

namespace zvx {

// ---------------------------------------------------------------------------------
// InstanceNorm statistics, per (utterance, channel) over the utterance's rows.
// Mirrors ggml_compute_forward_norm_f32 (/root/reference/ggml/src/ggml-cpu/ggml-cpu.c:
// 6905-6923): mean = (float)(sum_double / n); variance = (float)(sum_double((x-mean)^2) / n)
// with x-mean rounded to float first; scale = 1/sqrtf(variance + eps).
// grid (ceil(C/32), B), block (32, 8): a warp reads 32 consecutive channels of one row.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) stats_kernel(const float *__restrict__ x, int ld, int ch_off, int C,
                                                    const int *__restrict__ seg_start, int rate,
                                                    float *__restrict__ mu, float *__restrict__ rstd)
{
    __shared__ double red[8][33];
    __shared__ float mean_s[32];
    const int u  = blockIdx.y;
    const int c  = blockIdx.x * 32 + threadIdx.x;
    const int ty = threadIdx.y;
    const size_t r0 = (size_t)seg_start[u] * rate;
    const size_t r1 = (size_t)seg_start[u + 1] * rate;
    const double n = (double)(r1 - r0);
    const bool ok = c < C;

    double s = 0.0;
    if (ok)
        for (size_t r = r0 + ty; r < r1; r += 8) s += (double)x[r * ld + ch_off + c];
    red[ty][threadIdx.x] = s;
    __syncthreads();
    if (ty == 0) {
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
        mean_s[threadIdx.x] = (float)(t / n);
    }
    __syncthreads();
    const float mean = mean_s[threadIdx.x];
    double s2 = 0.0;
    if (ok)
        for (size_t r = r0 + ty; r < r1; r += 8) {
            const float v = __fsub_rn(x[r * ld + ch_off + c], mean);
            s2 += (double)__fmul_rn(v, v);
        }
    __syncthreads();
    red[ty][threadIdx.x] = s2;
    __syncthreads();
    if (ty == 0 && ok) {
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
        const float variance = (float)(t / n);
        mu[(size_t)u * C + c]   = mean;
        rstd[(size_t)u * C + c] = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(variance, 1e-5f)));
    }
}

cudaError_t stats_launch(const float *x, int ld, int ch_off, int C, const int *seg_start, int B, int rate, float *mu,
                         float *rstd, cudaStream_t st)
{
    dim3 grid((C + 31) / 32, B), block(32, 8);
    stats_kernel<<<grid, block, 0, st>>>(x, ld, ch_off, C, seg_start, rate, mu, rstd);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------
// AdaIN style projection for all AdaIN layers at once:  h = fc_w . s + fc_b;
// gamma1 = 1 + h[:C]; beta = h[C:]   (/root/reference/src/stylettsdec.cpp:177-189).
// One warp per output row n of the concatenated (sum 2C) x style_dim weight; the row is
// kept in registers and reused for every utterance of the batch.
// ---------------------------------------------------------------------------------
constexpr int ADAIN_MAX_S = 640;

__global__ void __launch_bounds__(256) adain_fc_kernel(const AdainTable tab, const float *__restrict__ style, int B,
                                                       float *__restrict__ out)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= tab.total) return;
    int k = 0, base = 0;
    while (k < tab.n - 1 && warp >= base + 2 * tab.d[k].C) {
        base += 2 * tab.d[k].C;
        ++k;
    }
    const AdainDesc d = tab.d[k];
    const int n = warp - base;                    // row inside this fc: [0, 2C)
    const int S = tab.style_dim;
    const float *wrow = d.fc_w + (size_t)n * S;
    float w[ADAIN_MAX_S / 32];                    // style_dim <= ADAIN_MAX_S
#pragma unroll
    for (int q = 0; q < ADAIN_MAX_S / 32; ++q) {
        const int i = lane + 32 * q;
        w[q] = i < S ? __ldg(wrow + i) : 0.f;
    }
    const float bias = __ldg(d.fc_b + n);
    for (int u = 0; u < B; ++u) {
        const float *s = style + (size_t)u * S;
        float acc = 0.f;
#pragma unroll
        for (int q = 0; q < ADAIN_MAX_S / 32; ++q) {
            const int i = lane + 32 * q;
            if (i < S) acc = fmaf(w[q], __ldg(s + i), acc);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) {
            float h = __fadd_rn(acc, bias);
            if (n < d.C) h = __fadd_rn(h, 1.0f);  // gamma + one (stylettsdec.cpp:189)
            out[(size_t)u * tab.total + d.out_off + n] = h;
        }
    }
}

cudaError_t adain_fc_launch(const AdainTable &tab, const float *style, int B, float *out, cudaStream_t st)
{
    if (tab.style_dim > ADAIN_MAX_S) return cudaErrorInvalidValue;
    const int warps_per_block = 8;
    const int blocks = (tab.total + warps_per_block - 1) / warps_per_block;
    adain_fc_kernel<<<blocks, warps_per_block * 32, 0, st>>>(tab, style, B, out);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------
// y = ((x - mu) * rstd) * w + b  (InstanceNorm1d affine of the asr_res branch,
// /root/reference/src/stylettsdec.cpp:392-396), written into the channel slice
// [dst_ch_off, dst_ch_off + C) of up to two concat buffers.
// grid (B, ceil(maxrows/ROWS_PER_BLOCK)) flattened as x: utterance-major loop inside.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) norm_affine_kernel(const float *__restrict__ x, int ldx, int C,
                                                          const int *__restrict__ seg_start, int B,
                                                          const float *__restrict__ mu, const float *__restrict__ rstd,
                                                          const float *__restrict__ w, const float *__restrict__ b,
                                                          float *__restrict__ dst0, float *__restrict__ dst1, int ld_dst,
                                                          int dst_ch_off)
{
    const int u = blockIdx.y;
    const size_t r0 = seg_start[u], r1 = seg_start[u + 1];
    const size_t total = (r1 - r0) * (size_t)C;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = r0 + i / C;
        const int c = (int)(i % C);
        float v = __fsub_rn(x[r * ldx + c], mu[(size_t)u * C + c]);
        v = __fmul_rn(v, rstd[(size_t)u * C + c]);
        v = __fmul_rn(v, __ldg(w + c));
        v = __fadd_rn(v, __ldg(b + c));
        dst0[r * ld_dst + dst_ch_off + c] = v;
        if (dst1) dst1[r * ld_dst + dst_ch_off + c] = v;
    }
}

cudaError_t norm_affine_launch(const float *x, int ldx, int C, const int *seg_start, int B, const float *mu,
                               const float *rstd, const float *w, const float *b, float *dst0, float *dst1, int ld_dst,
                               int dst_ch_off, cudaStream_t st)
{
    dim3 grid(32, B);
    norm_affine_kernel<<<grid, 256, 0, st>>>(x, ldx, C, seg_start, B, mu, rstd, w, b, dst0, dst1, ld_dst, dst_ch_off);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------
// Output conv of the vocoder: leaky_relu(0.01) -> Conv1d(C -> 1, K, pad (K-1)/2) + b -> tanh
// (/root/reference/src/hifigan.cpp:324-345).  One output channel: a dot product per sample
// on the CUDA cores.  A block stages (128 + K - 1) activated, fp16-rounded rows in shared
// memory (row stride C+1 floats: conflict-free), one thread per output sample.
// ---------------------------------------------------------------------------------
constexpr int OC_MAX_C = 64;
constexpr int OC_MAX_K = 16;

__global__ void __launch_bounds__(128) out_conv_kernel(const float *__restrict__ x, int C, int K,
                                                       const __half *__restrict__ w_raw, const float *__restrict__ bias,
                                                       float slope, const int *__restrict__ seg_start,
                                                       const int *__restrict__ tile_start, int B, int rate,
                                                       float *__restrict__ wav)
{
    extern __shared__ float sm[];
    float *ws   = sm;                           // [K][C]
    float *tile = sm + OC_MAX_K * OC_MAX_C;     // [(128 + K - 1)][C + 1]
    const int tile_id = blockIdx.x;
    const int u       = find_segment(tile_start, B, tile_id);
    const int t0      = (tile_id - tile_start[u]) * 128;
    const size_t row0 = (size_t)seg_start[u] * rate;
    const int seg_len = (seg_start[u + 1] - seg_start[u]) * rate;
    const int pad = (K - 1) / 2;
    const int rows = 128 + K - 1;

    for (int i = threadIdx.x; i < K * C; i += 128) {
        const int k = i / C, c = i % C;
        ws[k * C + c] = __half2float(w_raw[(size_t)c * K + k]);      // raw (OC=1, IC, K)
    }
    for (int i = threadIdx.x; i < rows * C; i += 128) {
        const int r = i / C, c = i % C;
        const int t = t0 - pad + r;
        float v = 0.f;
        if (t >= 0 && t < seg_len) v = __half2float(__float2half_rn(lrelu_f(x[(row0 + t) * C + c], slope)));
        tile[r * (C + 1) + c] = v;
    }
    __syncthreads();
    const int t = t0 + threadIdx.x;
    if (t >= seg_len) return;
    float acc = 0.f;
    for (int k = 0; k < K; ++k) {
        const float *row = tile + (threadIdx.x + k) * (C + 1);
        const float *wk = ws + k * C;
        for (int c = 0; c < C; ++c) acc = fmaf(row[c], wk[c], acc);
    }
    wav[row0 + t] = tanhf(__fadd_rn(acc, __ldg(bias)));
}

cudaError_t out_conv_launch(const float *x, int C, int K, const __half *w_raw, const float *bias, float slope,
                            const int *seg_start, const int *tile_start, int B, int rate, int total_tiles, float *wav,
                            cudaStream_t st)
{
    if (C > OC_MAX_C || K > OC_MAX_K) return cudaErrorInvalidValue;
    const size_t smem = (OC_MAX_K * OC_MAX_C + (128 + OC_MAX_K) * (OC_MAX_C + 1)) * sizeof(float);
    out_conv_kernel<<<total_tiles, 128, smem, st>>>(x, C, K, w_raw, bias, slope, seg_start, tile_start, B, rate, wav);
    return cudaGetLastError();
}

}  // namespace zvx
