// aux_kernels.cu -- the non-GEMM kernels of the hot path: InstanceNorm statistics, AdaIN
// style projection, affine-norm apply (asr_res branch) and the single-channel output conv.
// All are HBM/L2-bound; they use coalesced channel-contiguous accesses and shuffle / smem
// reductions.
#include "zvx_common.cuh"
#include "zvx_internal.h"

#include <algorithm>

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
    // One pass: sum and sum of squares in double.  The reference makes two passes (mean first,
    // then sum((float)(x - mean)^2) in double); the two variances differ by O(1e-7) relative
    // (the fp32 rounding of x - mean), far below the fp16 operand rounding that follows.
    __shared__ double red[2][8][33];
    const int u  = blockIdx.y;
    const int c  = blockIdx.x * 32 + threadIdx.x;
    const int ty = threadIdx.y;
    const size_t r0 = (size_t)seg_start[u] * rate;
    const size_t r1 = (size_t)seg_start[u + 1] * rate;
    const double n = (double)(r1 - r0);
    const bool ok = c < C;

    double s = 0.0, s2 = 0.0;
    if (ok) {
        const float *px = x + ch_off + c;
        size_t r = r0 + ty;
        for (; r + 24 < r1; r += 32) {             // 4 independent loads in flight per thread
            const float a0 = px[r * ld], a1 = px[(r + 8) * ld], a2 = px[(r + 16) * ld], a3 = px[(r + 24) * ld];
            s += (double)a0 + (double)a1 + (double)a2 + (double)a3;
            s2 += (double)a0 * a0 + (double)a1 * a1 + (double)a2 * a2 + (double)a3 * a3;
        }
        for (; r < r1; r += 8) {
            const float a0 = px[r * ld];
            s += (double)a0;
            s2 += (double)a0 * a0;
        }
    }
    red[0][ty][threadIdx.x] = s;
    red[1][ty][threadIdx.x] = s2;
    __syncthreads();
    if (ty == 0 && ok) {
        double t = 0.0, t2 = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) { t += red[0][i][threadIdx.x]; t2 += red[1][i][threadIdx.x]; }
        const double mean_d = t / n;
        double var_d = t2 / n - mean_d * mean_d;
        if (var_d < 0.0) var_d = 0.0;
        const float variance = (float)var_d;
        mu[(size_t)u * C + c]   = (float)mean_d;
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
// Second half of the statistics that the conv epilogues start (ConvParams::stats_out): per (utterance, channel) the
// per-tile (sum, sum of squares) pairs are added up in tile order -> mean / rstd exactly like stats_kernel.  Channels
// [C, C + C_tail) are copied from tail_mu / tail_rstd ([B][C_tail]; the asr_res part of the concatenated decoder input,
// whose statistics are computed once).  grid (ceil((C + C_tail) / 128), B).
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) stats_finalize_kernel(const double2 *__restrict__ part, int C, const int *__restrict__ tile_start,
                                                             const int *__restrict__ seg_start, int rate, const float *__restrict__ tail_mu,
                                                             const float *__restrict__ tail_rstd, int C_tail, float *__restrict__ mu,
                                                             float *__restrict__ rstd)
{
    const int u = blockIdx.y;
    const int c = blockIdx.x * 128 + threadIdx.x;
    const int stride = C + C_tail;
    if (c >= stride) return;
    if (c >= C) {
        mu[(size_t)u * stride + c]   = tail_mu[(size_t)u * C_tail + (c - C)];
        rstd[(size_t)u * stride + c] = tail_rstd[(size_t)u * C_tail + (c - C)];
        return;
    }
    const int t0 = tile_start[u], t1 = tile_start[u + 1];
    double s = 0.0, s2 = 0.0;
    for (int t = t0; t < t1; ++t) {
        const double2 v = part[(size_t)t * C + c];
        s += v.x;
        s2 += v.y;
    }
    const double n = (double)(seg_start[u + 1] - seg_start[u]) * rate;
    const double mean_d = s / n;
    double var_d = s2 / n - mean_d * mean_d;
    if (var_d < 0.0) var_d = 0.0;
    mu[(size_t)u * stride + c]   = (float)mean_d;
    rstd[(size_t)u * stride + c] = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn((float)var_d, 1e-5f)));
}

cudaError_t stats_finalize_launch(const double2 *part, int C, const int *tile_start, const int *seg_start, int B, int rate,
                                  const float *tail_mu, const float *tail_rstd, int C_tail, float *mu, float *rstd, cudaStream_t st)
{
    dim3 grid((C + C_tail + 127) / 128, B);
    stats_finalize_kernel<<<grid, 128, 0, st>>>(part, C, tile_start, seg_start, rate, tail_mu, tail_rstd, C_tail, mu, rstd);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------
// AdaIN style projection for all AdaIN layers at once:  h = fc_w . s + fc_b;
// gamma1 = 1 + h[:C]; beta = h[C:]   (/root/reference/src/stylettsdec.cpp:177-189).
// A small tiled fp32 GEMM: one block = 32 rows of the concatenated (sum 2C) x style_dim weight x 32 utterances, K in
// chunks of 64 through shared memory; a warp owns 4 rows (read as broadcasts), a lane one utterance.  (Round 1: one warp
// per row with a shuffle reduction per (row, utterance) -- 0.19 ms per step for 1.1 GFLOP.)
// ---------------------------------------------------------------------------------
constexpr int ADAIN_MAX_S = 640;
constexpr int AF_ROWS = 32, AF_UT = 32, AF_KC = 64, AF_LD = AF_KC + 4;     // + 4 floats: rows stay 16-byte aligned, quarter-warps conflict-free

__global__ void __launch_bounds__(256) adain_fc_kernel(const AdainTable tab, const float *__restrict__ style, int B,
                                                       float *__restrict__ out)
{
    __shared__ __align__(16) float ws[AF_ROWS][AF_LD];
    __shared__ __align__(16) float ss[AF_UT][AF_LD];
    __shared__ const float *rowp[AF_ROWS];
    __shared__ float rbias[AF_ROWS], radd[AF_ROWS];
    __shared__ int rout[AF_ROWS];
    const int tid = threadIdx.x, lane = tid & 31, rg = tid >> 5;
    const int S = tab.style_dim;
    if (tid < AF_ROWS) {
        const int ng = blockIdx.x * AF_ROWS + tid;
        rowp[tid] = nullptr;
        rout[tid] = -1;
        rbias[tid] = 0.f;
        radd[tid] = 0.f;
        if (ng < tab.total) {
            int k = 0, base = 0;
            while (k < tab.n - 1 && ng >= base + 2 * tab.d[k].C) {
                base += 2 * tab.d[k].C;
                ++k;
            }
            const AdainDesc d = tab.d[k];
            const int n = ng - base;                      // row inside this fc: [0, 2C)
            rowp[tid] = d.fc_w + (size_t)n * S;
            rbias[tid] = __ldg(d.fc_b + n);
            radd[tid] = n < d.C ? 1.0f : 0.0f;            // gamma + one (stylettsdec.cpp:189)
            rout[tid] = d.out_off + n;
        }
    }
    __syncthreads();
    for (int u0 = 0; u0 < B; u0 += AF_UT) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int k0 = 0; k0 < S; k0 += AF_KC) {
            for (int i = tid; i < AF_ROWS * AF_KC; i += 256) {
                const int r = i / AF_KC, c = i % AF_KC;
                ws[r][c] = (rowp[r] && k0 + c < S) ? __ldg(rowp[r] + k0 + c) : 0.f;
            }
            for (int i = tid; i < AF_UT * AF_KC; i += 256) {
                const int u = i / AF_KC, c = i % AF_KC;
                ss[u][c] = (u0 + u < B && k0 + c < S) ? __ldg(style + (size_t)(u0 + u) * S + k0 + c) : 0.f;
            }
            __syncthreads();
#pragma unroll 4
            for (int c = 0; c < AF_KC; c += 4) {
                const float4 sv = *reinterpret_cast<const float4 *>(&ss[lane][c]);
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const float4 wv = *reinterpret_cast<const float4 *>(&ws[rg * 4 + r][c]);
                    acc[r] = fmaf(wv.x, sv.x, acc[r]);
                    acc[r] = fmaf(wv.y, sv.y, acc[r]);
                    acc[r] = fmaf(wv.z, sv.z, acc[r]);
                    acc[r] = fmaf(wv.w, sv.w, acc[r]);
                }
            }
            __syncthreads();
        }
        if (u0 + lane < B) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int row = rg * 4 + r;
                if (rout[row] >= 0) {
                    float h = __fadd_rn(acc[r], rbias[row]);
                    if (radd[row] != 0.f) h = __fadd_rn(h, 1.0f);
                    out[(size_t)(u0 + lane) * tab.total + rout[row]] = h;
                }
            }
        }
    }
}

cudaError_t adain_fc_launch(const AdainTable &tab, const float *style, int B, float *out, cudaStream_t st)
{
    if (tab.style_dim > ADAIN_MAX_S) return cudaErrorInvalidValue;
    const int blocks = (tab.total + AF_ROWS - 1) / AF_ROWS;
    adain_fc_kernel<<<blocks, 256, 0, st>>>(tab, style, B, out);
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
// InstanceNorm / AdaIN apply + leaky-ReLU + fp16 rounding as a stand-alone pass:
//   y16 = fp16(lrelu(((x - mu) * rstd) * g + b, slope))      (same arithmetic and rounding order as
// the fused PRO_NORM prologue of conv_umma.cu; stylettsdec.cpp:94-104,191-197,253).
// The decoder convs split their output channels over 3-6 CTAs per time tile; with the prologue
// fused each of them re-did this math on the same rows, which made the A-operand producers the
// bottleneck (profiles/).  One pass here (10 B of HBM traffic per element) lets the convs load
// ready-made fp16 operands.
// ---------------------------------------------------------------------------------
constexpr int NA_ROWS = 8;       // rows per block (small: a half batch is only ~30 x 25 chunks of 32 rows, too few blocks to fill 148 SMs)

// grid (ceil(max_len / NA_ROWS), B), block = C/8 threads (one 8-channel group each, <= 256): a thread
// keeps its group's mean / rstd / gain / shift in registers and streams over the rows of its chunk;
// a warp reads / writes consecutive 32-byte / 16-byte pieces of one row.
__global__ void __launch_bounds__(256) norm_act_f16_kernel(const float *__restrict__ x, int ldx, int ch_off, int C,
                                                           const int *__restrict__ seg_start,
                                                           const float *__restrict__ mu, const float *__restrict__ rstd,
                                                           const float *__restrict__ g, const float *__restrict__ b,
                                                           int gb_stride, float slope, __half *__restrict__ y16,
                                                           __half *__restrict__ raw16)
{
    const int u = blockIdx.y;
    const size_t r0 = (size_t)seg_start[u] + (size_t)blockIdx.x * NA_ROWS;
    const size_t r1 = min(r0 + NA_ROWS, (size_t)seg_start[u + 1]);
    const int c = threadIdx.x * 8;
    if (r0 >= r1 || c >= C) return;
    ProCh pc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        pc[q].mu = __ldg(mu + (size_t)u * C + c + q);
        pc[q].rstd = __ldg(rstd + (size_t)u * C + c + q);
        pc[q].g = __ldg(g + (size_t)u * gb_stride + c + q);
        pc[q].b = __ldg(b + (size_t)u * gb_stride + c + q);
    }
    const float *px = x + ch_off + c;
    for (size_t r = r0; r < r1; r += 4) {
        float4 xa[4], xb[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const size_t rr = r + k < r1 ? r + k : r1 - 1;
            xa[k] = *reinterpret_cast<const float4 *>(px + rr * ldx);
            xb[k] = *reinterpret_cast<const float4 *>(px + rr * ldx + 4);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (r + k >= r1) break;
            const float xv[8] = {xa[k].x, xa[k].y, xa[k].z, xa[k].w, xb[k].x, xb[k].y, xb[k].z, xb[k].w};
            __half h[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) h[q] = __float2half_rn(prologue_apply(PRO_NORM, xv[q], slope, pc[q]));
            *reinterpret_cast<uint4 *>(y16 + (r + k) * C + c) = *reinterpret_cast<const uint4 *>(h);
            if (raw16) {        // the plain fp16 copy of the same rows (operand of the block's folded 1x1 shortcut): saves its own pass
#pragma unroll
                for (int q = 0; q < 8; ++q) h[q] = __float2half_rn(xv[q]);
                *reinterpret_cast<uint4 *>(raw16 + (r + k) * C + c) = *reinterpret_cast<const uint4 *>(h);
            }
        }
    }
}

// fp32 -> fp16 copy of an [rows][C] slice (operand of the 1x1 shortcut convs, which take the raw input)
__global__ void __launch_bounds__(256) cvt_f16_kernel(const float *__restrict__ x, int ldx, int ch_off, int C, size_t rows,
                                                      __half *__restrict__ y16)
{
    const int G = C >> 3;
    const size_t total = rows * (size_t)G;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t row = i / G;
        const int c = (int)(i % G) * 8;
        const float4 a = *reinterpret_cast<const float4 *>(x + row * ldx + ch_off + c);
        const float4 b = *reinterpret_cast<const float4 *>(x + row * ldx + ch_off + c + 4);
        __half h[8] = {__float2half_rn(a.x), __float2half_rn(a.y), __float2half_rn(a.z), __float2half_rn(a.w),
                       __float2half_rn(b.x), __float2half_rn(b.y), __float2half_rn(b.z), __float2half_rn(b.w)};
        *reinterpret_cast<uint4 *>(y16 + row * C + c) = *reinterpret_cast<const uint4 *>(h);
    }
}

// fp16( lrelu( ((a + b) + c) * scale ) ) of three [n] fp32 arrays: the MRF branch average of
// hifigan.cpp:300-315 followed by the leaky_relu of :281, materialised once as the fp16 operand of an
// up-conv whose output phases are separate launches (each of them would redo it otherwise).
__global__ void __launch_bounds__(256) sum3_act_f16_kernel(const float4 *__restrict__ a, const float4 *__restrict__ b,
                                                           const float4 *__restrict__ c, float scale, float slope, size_t n4,
                                                           uint2 *__restrict__ y16)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 va = a[i], vb = b[i], vc = c[i];
        const float v[4] = {__fmul_rn(__fadd_rn(__fadd_rn(va.x, vb.x), vc.x), scale), __fmul_rn(__fadd_rn(__fadd_rn(va.y, vb.y), vc.y), scale),
                            __fmul_rn(__fadd_rn(__fadd_rn(va.z, vb.z), vc.z), scale), __fmul_rn(__fadd_rn(__fadd_rn(va.w, vb.w), vc.w), scale)};
        const __half2 h01 = __floats2half2_rn(lrelu_f(v[0], slope), lrelu_f(v[1], slope));
        const __half2 h23 = __floats2half2_rn(lrelu_f(v[2], slope), lrelu_f(v[3], slope));
        uint2 h;
        h.x = *reinterpret_cast<const uint32_t *>(&h01);
        h.y = *reinterpret_cast<const uint32_t *>(&h23);
        y16[i] = h;
    }
}

// the same for fp16 branch tensors (8 elements per thread and step)
__global__ void __launch_bounds__(256) sum3h_act_f16_kernel(const uint4 *__restrict__ a, const uint4 *__restrict__ b,
                                                            const uint4 *__restrict__ c, float scale, float slope, size_t n8,
                                                            uint4 *__restrict__ y16)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 va = a[i], vb = b[i], vc = c[i];
        const uint32_t wa[4] = {va.x, va.y, va.z, va.w}, wb[4] = {vb.x, vb.y, vb.z, vb.w}, wc[4] = {vc.x, vc.y, vc.z, vc.w};
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 fa = __half22float2(*reinterpret_cast<const __half2 *>(&wa[k]));
            const float2 fb = __half22float2(*reinterpret_cast<const __half2 *>(&wb[k]));
            const float2 fc = __half22float2(*reinterpret_cast<const __half2 *>(&wc[k]));
            const __half2 h = __floats2half2_rn(lrelu_f(__fmul_rn(__fadd_rn(__fadd_rn(fa.x, fb.x), fc.x), scale), slope),
                                                lrelu_f(__fmul_rn(__fadd_rn(__fadd_rn(fa.y, fb.y), fc.y), scale), slope));
            o[k] = *reinterpret_cast<const uint32_t *>(&h);
        }
        y16[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

cudaError_t sum3h_act_f16_launch(const __half *a, const __half *b, const __half *c, float scale, float slope, size_t n, __half *y16,
                                 cudaStream_t st)
{
    if (n % 8) return cudaErrorInvalidValue;
    const size_t n8 = n / 8;
    int blocks = (int)std::min<size_t>((n8 + 255) / 256, (size_t)148 * 8);
    if (blocks < 1) blocks = 1;
    sum3h_act_f16_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const uint4 *>(a), reinterpret_cast<const uint4 *>(b),
                                                reinterpret_cast<const uint4 *>(c), scale, slope, n8, reinterpret_cast<uint4 *>(y16));
    return cudaGetLastError();
}

cudaError_t sum3_act_f16_launch(const float *a, const float *b, const float *c, float scale, float slope, size_t n, __half *y16,
                                cudaStream_t st)
{
    if (n % 4) return cudaErrorInvalidValue;
    const size_t n4 = n / 4;
    int blocks = (int)std::min<size_t>((n4 + 255) / 256, (size_t)148 * 8);
    if (blocks < 1) blocks = 1;
    sum3_act_f16_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float4 *>(a), reinterpret_cast<const float4 *>(b),
                                               reinterpret_cast<const float4 *>(c), scale, slope, n4, reinterpret_cast<uint2 *>(y16));
    return cudaGetLastError();
}

// Length regulator (the host loop at the end of FS2Encoder::eval, /root/reference/src/fs2encoder.cpp:611-641):
// the feature row of phoneme i is repeated tab[i].y times, starting at packed frame tab[i].x.  One CTA per
// phoneme: the row is read once, every copy is a run of full-line float4 stores.  Index work only -- the
// durations are rounded on the host with the reference's own libm call (zvx_api.cu).
__global__ void __launch_bounds__(128) length_regulate_kernel(const float4 *__restrict__ feat, const int2 *__restrict__ tab, int D4,
                                                              float4 *__restrict__ out)
{
    const int i = blockIdx.x;
    const int2 t = tab[i];
    for (int c = threadIdx.x; c < D4; c += 128) {
        const float4 v = feat[(size_t)i * D4 + c];
        for (int r = 0; r < t.y; ++r) out[(size_t)(t.x + r) * D4 + c] = v;
    }
}

cudaError_t length_regulate_launch(const float *feat, const int2 *tab, int n_phonemes, int D, float *out, cudaStream_t st)
{
    if (D % 4) return cudaErrorInvalidValue;
    if (n_phonemes <= 0) return cudaSuccess;
    length_regulate_kernel<<<n_phonemes, 128, 0, st>>>(reinterpret_cast<const float4 *>(feat), tab, D / 4, reinterpret_cast<float4 *>(out));
    return cudaGetLastError();
}

cudaError_t cvt_f16_launch(const float *x, int ldx, int ch_off, int C, size_t rows, __half *y16, cudaStream_t st)
{
    if (C % 8) return cudaErrorInvalidValue;
    const size_t total = rows * (size_t)(C / 8);
    int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)148 * 8);
    if (blocks < 1) blocks = 1;
    cvt_f16_kernel<<<blocks, 256, 0, st>>>(x, ldx, ch_off, C, rows, y16);
    return cudaGetLastError();
}

cudaError_t norm_act_f16_launch(const float *x, int ldx, int ch_off, int C, const int *seg_start, int B, int max_len,
                                const float *mu, const float *rstd, const float *g, const float *b, int gb_stride, float slope,
                                __half *y16, __half *raw16, cudaStream_t st)
{
    if (C % 8 || C / 8 > 256) return cudaErrorInvalidValue;
    dim3 grid((max_len + NA_ROWS - 1) / NA_ROWS, B);
    const int threads = ((C / 8 + 31) / 32) * 32;
    norm_act_f16_kernel<<<grid, threads, 0, st>>>(x, ldx, ch_off, C, seg_start, mu, rstd, g, b, gb_stride, slope, y16, raw16);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------
// Output conv of the vocoder: leaky_relu(0.01) -> Conv1d(C -> 1, K, pad (K-1)/2) + b -> tanh
// (/root/reference/src/hifigan.cpp:324-345).  One output channel: a dot product per sample
// on the CUDA cores.  A block stages (128 + K - 1) activated, fp16-rounded rows in shared
// memory (row stride C+1 floats: conflict-free), one thread per output sample.
// ---------------------------------------------------------------------------------
// One output sample: float (what HiFiGAN::eval returns, hifigan.cpp:374-376) and / or signed 16-bit PCM, i.e. the
// conversion libsndfile applies inside sf_write_float() for SF_FORMAT_PCM_16 (the reference's write_wav_file,
// zerovox.cpp:357-371; libsndfile src/pcm.c f2s_array with the default SFC_SET_NORM_FLOAT = true and clipping
// off: lrintf(x * 0x7FFF), round-half-even).  |tanh| <= 1, so the product is always in range.
__device__ __forceinline__ void store_sample(float v, size_t i, float *__restrict__ wav, int16_t *__restrict__ pcm)
{
    if (wav) wav[i] = v;
    if (pcm) pcm[i] = (int16_t)__float2int_rn(__fmul_rn(v, 32767.0f));
}

constexpr int OC_MAX_C = 64;
constexpr int OC_MAX_K = 16;

// Generic shape: weights staged in shared memory.
__global__ void __launch_bounds__(128) out_conv_kernel(const float *__restrict__ x, const float *__restrict__ x2,
                                                       const float *__restrict__ x3, const __half *__restrict__ x16, const int halves, float sum_scale, int C, int K,
                                                       const __half *__restrict__ w_raw, const float *__restrict__ bias,
                                                       float slope, const int *__restrict__ seg_start,
                                                       const int *__restrict__ tile_start, int B, int rate,
                                                       float *__restrict__ wav, int16_t *__restrict__ pcm)
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
        if (t >= 0 && t < seg_len) {
            if (x16) {
                v = __half2float(x16[(row0 + t) * C + c]);          // already averaged, activated and rounded by the producer
            } else {
                float xv;
                if (halves) {       // the three branch tensors arrive as fp16
                    const size_t e = (row0 + t) * C + c;
                    xv = __fmul_rn(__fadd_rn(__fadd_rn(__half2float(reinterpret_cast<const __half *>(x)[e]), __half2float(reinterpret_cast<const __half *>(x2)[e])),
                                             __half2float(reinterpret_cast<const __half *>(x3)[e])), sum_scale);
                } else {
                    xv = x[(row0 + t) * C + c];
                    if (x2) xv = __fmul_rn(__fadd_rn(__fadd_rn(xv, x2[(row0 + t) * C + c]), x3[(row0 + t) * C + c]), sum_scale);
                }
                v = __half2float(__float2half_rn(lrelu_f(xv, slope)));
            }
        }
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
    store_sample(tanhf(__fadd_rn(acc, __ldg(bias))), row0 + t, wav, pcm);
}

// The shipped shape (32 channels, 7 taps): the 224 weights travel as a kernel parameter, i.e. in the
// constant bank, so every FFMA takes its weight as an immediate constant operand and the inner loop
// is one conflict-free LDS + one FFMA per MAC; the input tile is staged with float4 loads.
struct OutConvW { float w[7 * 32]; float bias; };      // [k][c]

__global__ void __launch_bounds__(128) out_conv_32x7_kernel(const float *__restrict__ x, const float *__restrict__ x2,
                                                            const float *__restrict__ x3, const __half *__restrict__ x16, const int halves, float sum_scale,
                                                            const OutConvW W, float slope,
                                                            const int *__restrict__ seg_start,
                                                            const int *__restrict__ tile_start, int B, int rate,
                                                            float *__restrict__ wav, int16_t *__restrict__ pcm)
{
    constexpr int C = 32, K = 7, ROWS = 128 + K - 1, LD = C + 1;
    __shared__ float tile[ROWS * LD];
    const int tile_id = blockIdx.x;
    const int u       = find_segment(tile_start, B, tile_id);
    const int t0      = (tile_id - __ldg(tile_start + u)) * 128;
    const size_t row0 = (size_t)__ldg(seg_start + u) * rate;
    const int seg_len = (__ldg(seg_start + u + 1) - __ldg(seg_start + u)) * rate;

    for (int i = threadIdx.x; i < ROWS * (C / 4); i += 128) {
        const int r = i >> 3, c4 = (i & 7) * 4;
        const int t = t0 - (K - 1) / 2 + r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (x16) {
            // ready-made operand (the last residual block's final phase averaged the branches, applied the
            // leaky-ReLU and rounded to fp16): 8 bytes per 4 channels instead of 3 x 16
            float *d = tile + r * LD + c4;
            if (t >= 0 && t < seg_len) {
                const uint2 h = *reinterpret_cast<const uint2 *>(x16 + (row0 + t) * C + c4);
                const __half2 h01 = *reinterpret_cast<const __half2 *>(&h.x), h23 = *reinterpret_cast<const __half2 *>(&h.y);
                d[0] = __low2float(h01); d[1] = __high2float(h01); d[2] = __low2float(h23); d[3] = __high2float(h23);
            } else {
                d[0] = d[1] = d[2] = d[3] = 0.f;
            }
            continue;
        }
        if (halves) {
            // the three branch tensors arrive as fp16 (6 instead of 12 bytes per element)
            if (t >= 0 && t < seg_len) {
                const size_t e = (row0 + t) * C + c4;
                const uint2 qa = *reinterpret_cast<const uint2 *>(reinterpret_cast<const __half *>(x) + e);
                const uint2 qb = *reinterpret_cast<const uint2 *>(reinterpret_cast<const __half *>(x2) + e);
                const uint2 qc = *reinterpret_cast<const uint2 *>(reinterpret_cast<const __half *>(x3) + e);
                const float2 a0 = __half22float2(*reinterpret_cast<const __half2 *>(&qa.x)), a1 = __half22float2(*reinterpret_cast<const __half2 *>(&qa.y));
                const float2 b0 = __half22float2(*reinterpret_cast<const __half2 *>(&qb.x)), b1 = __half22float2(*reinterpret_cast<const __half2 *>(&qb.y));
                const float2 c0 = __half22float2(*reinterpret_cast<const __half2 *>(&qc.x)), c1 = __half22float2(*reinterpret_cast<const __half2 *>(&qc.y));
                v.x = __fmul_rn(__fadd_rn(__fadd_rn(a0.x, b0.x), c0.x), sum_scale);
                v.y = __fmul_rn(__fadd_rn(__fadd_rn(a0.y, b0.y), c0.y), sum_scale);
                v.z = __fmul_rn(__fadd_rn(__fadd_rn(a1.x, b1.x), c1.x), sum_scale);
                v.w = __fmul_rn(__fadd_rn(__fadd_rn(a1.y, b1.y), c1.y), sum_scale);
            }
        } else if (t >= 0 && t < seg_len) {
            v = *reinterpret_cast<const float4 *>(x + (row0 + t) * C + c4);
            if (x2) {
                // branch sum / average of the last MRF stage (hifigan.cpp:300-315), same order as the reference
                const float4 b = *reinterpret_cast<const float4 *>(x2 + (row0 + t) * C + c4);
                const float4 c = *reinterpret_cast<const float4 *>(x3 + (row0 + t) * C + c4);
                v.x = __fmul_rn(__fadd_rn(__fadd_rn(v.x, b.x), c.x), sum_scale);
                v.y = __fmul_rn(__fadd_rn(__fadd_rn(v.y, b.y), c.y), sum_scale);
                v.z = __fmul_rn(__fadd_rn(__fadd_rn(v.z, b.z), c.z), sum_scale);
                v.w = __fmul_rn(__fadd_rn(__fadd_rn(v.w, b.w), c.w), sum_scale);
            }
        }
        float *d = tile + r * LD + c4;
        d[0] = __half2float(__float2half_rn(lrelu_f(v.x, slope)));
        d[1] = __half2float(__float2half_rn(lrelu_f(v.y, slope)));
        d[2] = __half2float(__float2half_rn(lrelu_f(v.z, slope)));
        d[3] = __half2float(__float2half_rn(lrelu_f(v.w, slope)));
    }
    __syncthreads();
    const int t = t0 + threadIdx.x;
    if (t >= seg_len) return;
    float acc0 = 0.f, acc1 = 0.f;
    const float *row = tile + threadIdx.x * LD;
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
        for (int c = 0; c < C; c += 2) {
            acc0 = fmaf(row[k * LD + c], W.w[k * C + c], acc0);
            acc1 = fmaf(row[k * LD + c + 1], W.w[k * C + c + 1], acc1);
        }
    }
    store_sample(tanhf(__fadd_rn(__fadd_rn(acc0, acc1), W.bias)), row0 + t, wav, pcm);
}

cudaError_t out_conv_launch(const float *x, const float *x2, const float *x3, const __half *x16, int halves, float sum_scale, int C, int K, const __half *w_raw, const float *bias, const float *w_host_kc,
                            float bias_host, float slope, const int *seg_start, const int *tile_start, int B, int rate,
                            int total_tiles, float *wav, int16_t *pcm, cudaStream_t st)
{
    if (C == 32 && K == 7 && w_host_kc) {
        OutConvW W;
        for (int i = 0; i < 7 * 32; ++i) W.w[i] = w_host_kc[i];
        W.bias = bias_host;
        out_conv_32x7_kernel<<<total_tiles, 128, 0, st>>>(x, x2, x3, x16, halves, sum_scale, W, slope, seg_start, tile_start, B, rate, wav, pcm);
        return cudaGetLastError();
    }
    if (C > OC_MAX_C || K > OC_MAX_K) return cudaErrorInvalidValue;
    const size_t smem = (OC_MAX_K * OC_MAX_C + (128 + OC_MAX_K) * (OC_MAX_C + 1)) * sizeof(float);
    out_conv_kernel<<<total_tiles, 128, smem, st>>>(x, x2, x3, x16, halves, sum_scale, C, K, w_raw, bias, slope, seg_start, tile_start, B, rate, wav, pcm);
    return cudaGetLastError();
}

}  // namespace zvx
