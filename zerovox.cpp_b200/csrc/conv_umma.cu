// conv_umma.cu -- 1-D convolution as an implicit GEMM on the sm_100a tensor cores.
//
// Replaces, for every conv on the hot path, the reference sequence
//   ggml_conv_1d = IM2COL(F32->F16) + MUL_MAT(F16 x F16, fp32 accumulate)
//     /root/reference/ggml/src/ggml.c:3769-3786, ggml-cpu/ggml-cpu.c:9890-9961, :7377-7554
//   + cont(transpose) + add(repeat(bias)) + cont(transpose)     (e.g. hifigan.cpp:138-140)
// and the elementwise op that feeds it (leaky_relu / InstanceNorm-affine / AdaIN /
// mel normalisation) and follows it (bias, residual add, MRF branch sum, 1/sqrt2, 1/3).
//
// One CTA computes a [128 * MT time steps] x [NC output channels] tile (MT = 1 or 2 M-tiles that
// share every weight stage: with MT = 2 each weight byte fetched from L2 feeds 2 x 128 rows,
// halving the L2 -> smem weight traffic that bounds the MT = 1 kernel):
//     D[t, oc] = sum_taps sum_ic  A[t + off(tap), ic] * W[tap][oc, ic]
//   * A operand: warps 0-3 read a halo tile (128 + (ntaps-1)*dilation rows) of the
//     channels-last activation ONCE per 64-channel chunk, apply the fused prologue, round
//     to fp16 and store it to shared memory in the UMMA K-major no-swizzle ("interleave")
//     canonical layout with an 8-row-group stride of 128 B, i.e. row r of 8-channel group j
//     lives at j*LBO + r*16.  Rows are uniformly 16 B apart, so each tap is just the same
//     tile re-addressed through the matrix descriptor's start address (+ tap*dilation*16 B):
//     no im2col, no per-tap reload (SURVEY.md H-b).
//   * B operand: weights are pre-packed at load time into per-(N-chunk, K-chunk, tap)
//     blocks that are already in the same canonical layout; one thread streams them
//     into a ring of stages with cp.async.bulk (TMA bulk copy) completing on mbarriers.
//   * MMA: one thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=NC, K=16),
//     fp16 x fp16 -> fp32 accumulators in tensor memory; tcgen05.commit releases the
//     smem stages and finally signals the epilogue.
//   * Epilogue: warps 0-3 read the accumulator with tcgen05.ld (one row per thread) and
//     apply bias / residual / branch-sum / scale, writing fp32 and/or the next conv's
//     fp16 pre-activated operand.
#include <cstring>

#include <cuda.h>

#include "zvx_common.cuh"
#include "zvx_internal.h"
#include "ptx_sm100.cuh"

namespace zvx {

constexpr int TILE_M      = 128;
constexpr int KCHUNK      = 64;
constexpr int MAX_A_STAGES = 4;
constexpr int MAX_B_STAGES = 8;
constexpr int SMEM_HEADER  = 384;

__device__ __forceinline__ uint32_t make_idesc(int N) { return make_idesc_mn(TILE_M, N); }

// completion of this thread's earlier cp.async copies counts as its arrival on the barrier
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

// ------------------------------------------------------------------ shared pieces
struct TileCoord {
    int u;            // utterance
    int t0;           // first output row of the tile inside the utterance (input-rate rows)
    int seg_len;      // rows of the utterance
    size_t seg_row0;  // first row of the utterance in the packed tensors
    int nchunk;       // N-chunk (output channels [nchunk*NC, +NC))
};

// One 64-channel K-chunk of the A operand: read the halo tile, apply the fused prologue, round to
// fp16 and store it in the no-swizzle K-major layout (8-row groups 128 B apart, channel groups
// lbo_a apart).  `tid` is the producer-thread index in [0, NPROD).
template <int MODE, int NPROD, int UNR_>
__device__ __forceinline__ void produce_chunk(const ConvParams &p, const TileCoord &tc, int c, uint8_t *stage, uint32_t lbo_a,
                                              int need_rows, int tid)
{
    const int kc   = min(KCHUNK, p.Cin - c * KCHUNK);
    const int G    = kc >> 3;                 // 8-channel groups in this chunk: 8, 4 or 2
    const int lg   = (G == 8) ? 3 : (G == 4) ? 2 : 1;
    const int j    = tid & (G - 1);
    const int r0   = tid >> lg;
    const int rstep = NPROD >> lg;
    const int ch   = c * KCHUNK + j * 8;      // channel (relative to x_ch_off) of this thread's group

    ProCh pc[8];
    if (MODE == PRO_NORM || MODE == PRO_MEL) {
        const float *mu = p.p_mu + (size_t)tc.u * p.p_stat_stride + ch;
        const float *rs = p.p_rstd + (size_t)tc.u * p.p_stat_stride + ch;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            pc[i].mu = __ldg(mu + i);
            pc[i].rstd = __ldg(rs + i);
        }
        if (MODE == PRO_NORM) {
            const float *g = p.p_g + (size_t)tc.u * p.p_gb_stride + ch;
            const float *b = p.p_b + (size_t)tc.u * p.p_gb_stride + ch;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                pc[i].g = __ldg(g + i);
                pc[i].b = __ldg(b + i);
            }
        }
    }

    // Rows of this thread: rho = r0 + i * rstep.  Loads of a batch of UNR rows are issued
    // before any of them is used (memory-level parallelism: the producer is latency-bound
    // otherwise), then transformed and stored with plain shared-memory stores.
    uint8_t *dstp = stage + (size_t)j * lbo_a;
    if (MODE == PRO_F16) {
        // ready-made fp16 operand: asynchronous 16-byte copies global -> shared (zero-filled outside the
        // utterance); the caller makes the stage's mbarrier track their completion, so a producer
        // thread can have several K-chunks in flight instead of one batch of loads at a time
        // Batches of 4 copies with their own address registers: a rolled loop re-uses the registers of the previous
        // cp.async and stalls until that instruction has left the LSU queue (profiles/r02_ncu_conv_dec_before.txt: 43 %
        // of all warp samples of the decoder convs sat on that one dependency).
        const uint32_t dst0 = smem_u32(dstp);
        const __half *xh = reinterpret_cast<const __half *>(p.x) + (size_t)p.x_ch_off + ch;
        const size_t row0 = tc.seg_row0;
        const int tbase = tc.t0 + p.tap_off0;
        for (int rho0 = r0; rho0 < need_rows; rho0 += 4 * rstep) {
            const __half *src[4];
            uint32_t dst[4];
            int nbytes[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int rho = rho0 + q * rstep;
                const int t_in = tbase + rho;
                const bool ok = t_in >= 0 && t_in < tc.seg_len;
                src[q] = xh + (row0 + (size_t)(ok ? t_in : 0)) * (size_t)p.ldx;
                dst[q] = dst0 + (uint32_t)rho * 16u;
                nbytes[q] = ok ? 16 : 0;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (rho0 + q * rstep < need_rows)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst[q]), "l"(src[q]), "r"(nbytes[q]) : "memory");
        }
    } else if (MODE == PRO_SUM3H) {
        // three fp16 sources per element (branch outputs of the fused MRF blocks): 4 rows x 3 x 16 bytes in flight
        const __half *h1 = reinterpret_cast<const __half *>(p.x), *h2 = reinterpret_cast<const __half *>(p.x2),
                     *h3 = reinterpret_cast<const __half *>(p.x3);
        for (int rho0 = r0; rho0 < need_rows; rho0 += 4 * rstep) {
            uint4 q[4][3];
            bool ok[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int rho  = rho0 + k * rstep;
                const int t_in = tc.t0 + p.tap_off0 + rho;
                ok[k] = rho < need_rows && t_in >= 0 && t_in < tc.seg_len;
                const size_t e = (tc.seg_row0 + (size_t)(ok[k] ? t_in : 0)) * (size_t)p.ldx + p.x_ch_off + ch;
                if (ok[k]) {
                    q[k][0] = *reinterpret_cast<const uint4 *>(h1 + e);
                    q[k][1] = *reinterpret_cast<const uint4 *>(h2 + e);
                    q[k][2] = *reinterpret_cast<const uint4 *>(h3 + e);
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int rho = rho0 + k * rstep;
                if (rho >= need_rows) break;
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                if (ok[k]) {
                    const uint32_t wa[4] = {q[k][0].x, q[k][0].y, q[k][0].z, q[k][0].w};
                    const uint32_t wb[4] = {q[k][1].x, q[k][1].y, q[k][1].z, q[k][1].w};
                    const uint32_t wc[4] = {q[k][2].x, q[k][2].y, q[k][2].z, q[k][2].w};
                    uint32_t o[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&wa[i]));
                        const float2 b = __half22float2(*reinterpret_cast<const __half2 *>(&wb[i]));
                        const float2 c = __half22float2(*reinterpret_cast<const __half2 *>(&wc[i]));
                        o[i] = pack_half2(lrelu_f(__fmul_rn(__fadd_rn(__fadd_rn(a.x, b.x), c.x), p.sum_scale), p.pro_slope),
                                          lrelu_f(__fmul_rn(__fadd_rn(__fadd_rn(a.y, b.y), c.y), p.sum_scale), p.pro_slope));
                    }
                    v = make_uint4(o[0], o[1], o[2], o[3]);
                }
                *reinterpret_cast<uint4 *>(dstp + (size_t)rho * 16) = v;
            }
        }
    } else if (MODE == PRO_SUM3) {
        // three fp32 sources per element: smaller batches (2 rows x 3 sources x 2 float4 in flight)
        const float *x1 = reinterpret_cast<const float *>(p.x);
        for (int rho0 = r0; rho0 < need_rows; rho0 += 2 * rstep) {
            float4 f[2][6];
            bool ok[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int rho  = rho0 + q * rstep;
                const int t_in = tc.t0 + p.tap_off0 + rho;
                ok[q] = rho < need_rows && t_in >= 0 && t_in < tc.seg_len;
                const size_t e = (tc.seg_row0 + (size_t)(ok[q] ? t_in : 0)) * (size_t)p.ldx + p.x_ch_off + ch;
#pragma unroll
                for (int i = 0; i < 6; ++i) f[q][i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ok[q]) {
                    f[q][0] = *reinterpret_cast<const float4 *>(x1 + e);
                    f[q][1] = *reinterpret_cast<const float4 *>(x1 + e + 4);
                    f[q][2] = *reinterpret_cast<const float4 *>(p.x2 + e);
                    f[q][3] = *reinterpret_cast<const float4 *>(p.x2 + e + 4);
                    f[q][4] = *reinterpret_cast<const float4 *>(p.x3 + e);
                    f[q][5] = *reinterpret_cast<const float4 *>(p.x3 + e + 4);
                }
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int rho = rho0 + q * rstep;
                if (rho >= need_rows) break;
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                if (ok[q]) {
                    const float a[8] = {f[q][0].x, f[q][0].y, f[q][0].z, f[q][0].w, f[q][1].x, f[q][1].y, f[q][1].z, f[q][1].w};
                    const float b[8] = {f[q][2].x, f[q][2].y, f[q][2].z, f[q][2].w, f[q][3].x, f[q][3].y, f[q][3].z, f[q][3].w};
                    const float c[8] = {f[q][4].x, f[q][4].y, f[q][4].z, f[q][4].w, f[q][5].x, f[q][5].y, f[q][5].z, f[q][5].w};
                    float o[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        o[i] = lrelu_f(__fmul_rn(__fadd_rn(__fadd_rn(a[i], b[i]), c[i]), p.sum_scale), p.pro_slope);
                    v.x = pack_half2(o[0], o[1]);
                    v.y = pack_half2(o[2], o[3]);
                    v.z = pack_half2(o[4], o[5]);
                    v.w = pack_half2(o[6], o[7]);
                }
                *reinterpret_cast<uint4 *>(dstp + (size_t)rho * 16) = v;
            }
        }
    } else {
    constexpr int UNR = UNR_;
    for (int rho0 = r0; rho0 < need_rows; rho0 += UNR * rstep) {
        float4 fa[UNR], fb[UNR];
        uint4  hv[UNR];
        bool   ok[UNR];
#pragma unroll
        for (int q = 0; q < UNR; ++q) {
            const int rho  = rho0 + q * rstep;
            const int t_in = tc.t0 + p.tap_off0 + rho;
            ok[q] = rho < need_rows && t_in >= 0 && t_in < tc.seg_len;
            const size_t e = (tc.seg_row0 + (size_t)(ok[q] ? t_in : 0)) * (size_t)p.ldx + p.x_ch_off + ch;
            if (MODE == PRO_F16) {
                hv[q] = make_uint4(0u, 0u, 0u, 0u);
                if (ok[q]) hv[q] = *reinterpret_cast<const uint4 *>(reinterpret_cast<const __half *>(p.x) + e);
            } else {
                fa[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                fb[q] = fa[q];
                if (ok[q]) {
                    const float4 *src = reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(p.x) + e);
                    fa[q] = src[0];
                    fb[q] = src[1];
                }
            }
        }
#pragma unroll
        for (int q = 0; q < UNR; ++q) {
            const int rho = rho0 + q * rstep;
            if (rho >= need_rows) break;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (MODE == PRO_F16) {
                v = hv[q];
            } else if (ok[q]) {
                float f[8] = {fa[q].x, fa[q].y, fa[q].z, fa[q].w, fb[q].x, fb[q].y, fb[q].z, fb[q].w};
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] = prologue_apply(MODE, f[i], p.pro_slope, pc[i]);
                v.x = pack_half2(f[0], f[1]);
                v.y = pack_half2(f[2], f[3]);
                v.z = pack_half2(f[4], f[5]);
                v.w = pack_half2(f[6], f[7]);
            }
            *reinterpret_cast<uint4 *>(dstp + (size_t)rho * 16) = v;
        }
    }
    }
}

// Epilogue of one warp = 32 accumulator rows (time steps) x a share of the NC columns.  tcgen05.ld hands every thread
// one ROW; writing rows from there costs 32 cache lines per warp instruction.  So the warp transposes through a private
// shared-memory slab, 32 columns at a time: afterwards a quarter-warp owns 128 contiguous bytes of one row and every
// global access (residual / branch-sum reads, fp32 and fp16 writes) covers whole lines.
// Round 2 (profiles/r02_ncu_conv_*): this code, not the MMAs, bounded every MRF stage-0 launch and a third of a decoder
// conv's life -- ~1700 instructions per warp and 64-column pass (64-bit address arithmetic, per-element flag tests, a
// 5-instruction leaky-ReLU).  Now: row pointers advance by precomputed strides, the flags are hoisted, the residual of a
// pass is requested before its accumulator columns are pulled out of tensor memory, leaky-ReLU is max(x, a x) (equal to
// ggml's max(x,0) + a min(x,0) for 0 < a < 1 up to the sign of zero): ~220 instructions per 32-column pass.
// Fused math, unchanged: ((acc + bias) + residual) + acc_in, times scale; fp16 copy with leaky-ReLU.
constexpr int EPI_COLS   = 32;
constexpr int SLAB_LD    = EPI_COLS + 4;               // floats; +4 keeps the row-wise 16-byte stores conflict-free
constexpr int SLAB_BYTES = 32 * SLAB_LD * 4;

__device__ __forceinline__ float4 f4_add(float4 a, float4 b)
{
    return make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w));
}
__device__ __forceinline__ float lrelu_max_f(float x, float a) { return fmaxf(x, __fmul_rn(a, x)); }

// 32 columns x 32 rows of tensor memory -> registers (two x16 loads in flight, one wait)
__device__ __forceinline__ void tmem_ld16x2(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%32];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%33];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr), "r"(taddr + 16u));
}

// Columns are handled in passes of EPI_COLS; this warp takes the passes pass0, pass0 + pass_step, ... (the persistent
// kernel runs two warps per lane quarter).  trow: tensor-memory address of the warp's lane quarter, column 0 of the tile.
// stat_row (one-tile kernel, may be null): [NC] double2 of this warp's lane quarter, filled with the column sums / sums of
// squares of the final values over the warp's valid rows.
template <bool STATS>
__device__ __forceinline__ void epilogue_tile(const ConvParams &p, uint32_t trow, float *slab, int lane, int t_first, int seg_len,
                                              size_t seg_row0, int nchunk, int NC, int pass0, int pass_step, double2 *stat_row_ = nullptr)
{
    double2 *const stat_row = STATS ? stat_row_ : nullptr;      // (kernels whose outputs never feed a norm compile without the sums)
    const int cl   = (lane & 7) * 4;        // this lane's 4 columns inside the 32-column pass
    const int rsel = lane >> 3;             // which row of a group of 4
    const bool has_res = p.res != nullptr, has_acc = p.acc_in != nullptr, has_o32 = p.out32 != nullptr, has_o16 = p.out16 != nullptr;
    const float scale = p.has_scale ? p.scale : 1.0f;          // v * 1.0f == v
    const float slope = p.out16_slope;
    const int rows_valid = seg_len - t_first - rsel;            // iteration i (row 4 i + rsel) is inside the utterance iff 4 i < rows_valid
    // element offsets of row (t_first + rsel), advancing by 4 rows per iteration
    const size_t orow0 = (seg_row0 + (size_t)(t_first + rsel)) * (size_t)p.out_mul + p.out_add;
    const size_t rstep = (size_t)4 * p.out_mul;
    const float *res_row = has_res ? p.res + orow0 * (size_t)p.ldres + p.res_ch_off : nullptr;
    const float *acc_row = has_acc ? p.acc_in + orow0 * (size_t)p.ldo32 + p.o32_ch_off : nullptr;
    float *o32_row = has_o32 ? p.out32 + orow0 * (size_t)p.ldo32 + p.o32_ch_off : nullptr;
    __half *o16_row = has_o16 ? p.out16 + orow0 * (size_t)p.ldo16 + p.o16_ch_off : nullptr;
    const size_t res_step = rstep * (size_t)p.ldres, o32_step = rstep * (size_t)p.ldo32, o16_step = rstep * (size_t)p.ldo16;
    const float4 *srow = reinterpret_cast<const float4 *>(slab + rsel * SLAB_LD + cl);
    float4 *drow = reinterpret_cast<float4 *>(slab + lane * SLAB_LD);

    for (int col0 = pass0 * EPI_COLS; col0 < NC; col0 += pass_step * EPI_COLS) {
        const int cw = min(EPI_COLS, NC - col0);
        const int oc = nchunk * NC + col0 + cl;
        const bool col_ok = cl < cw;
        // residual rows of the pass first: their latency is covered by the tensor-memory read-out and the transpose
        float4 rs[8];
        if (has_res) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                rs[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (col_ok && 4 * i < rows_valid) rs[i] = *reinterpret_cast<const float4 *>(res_row + (size_t)i * res_step + oc);
            }
        }
        {
            uint32_t r[32];
            tmem_ld16x2(trow + (uint32_t)col0, r);
#pragma unroll
            for (int q = 0; q < 8; ++q)
                drow[q] = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]),
                                      __uint_as_float(r[4 * q + 3]));
        }
        __syncwarp();
        // statistics: this thread's (up to) 8 rows of a column are summed in fp32 (8 terms: relative error ~1e-7, far below
        // the fp16 rounding of the operand the normalised tensor becomes), everything across threads / tiles in double
        float fx[4] = {0.f, 0.f, 0.f, 0.f}, fq[4] = {0.f, 0.f, 0.f, 0.f};
        if (col_ok) {
            float4 bias = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.bias) bias = __ldg(reinterpret_cast<const float4 *>(p.bias + oc));
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (4 * i < rows_valid) {
                    float4 v = srow[i * SLAB_LD];
                    if (p.bias) v = f4_add(v, bias);
                    if (has_res) v = f4_add(v, rs[i]);
                    if (has_acc) v = f4_add(*reinterpret_cast<const float4 *>(acc_row + (size_t)i * o32_step + oc), v);
                    v = make_float4(__fmul_rn(v.x, scale), __fmul_rn(v.y, scale), __fmul_rn(v.z, scale), __fmul_rn(v.w, scale));
                    if (has_o32) *reinterpret_cast<float4 *>(o32_row + (size_t)i * o32_step + oc) = v;
                    if (stat_row) {
                        fx[0] = __fadd_rn(fx[0], v.x); fx[1] = __fadd_rn(fx[1], v.y); fx[2] = __fadd_rn(fx[2], v.z); fx[3] = __fadd_rn(fx[3], v.w);
                        fq[0] = __fmaf_rn(v.x, v.x, fq[0]); fq[1] = __fmaf_rn(v.y, v.y, fq[1]); fq[2] = __fmaf_rn(v.z, v.z, fq[2]);
                        fq[3] = __fmaf_rn(v.w, v.w, fq[3]);
                    }
                    if (has_o16) {
                        uint2 h;
                        h.x = pack_half2(lrelu_max_f(v.x, slope), lrelu_max_f(v.y, slope));
                        h.y = pack_half2(lrelu_max_f(v.z, slope), lrelu_max_f(v.w, slope));
                        *reinterpret_cast<uint2 *>(o16_row + (size_t)i * o16_step + oc) = h;
                    }
                }
            }
        }
        if (stat_row) {
            double sx[4], sq[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) { sx[k] = (double)fx[k]; sq[k] = (double)fq[k]; }
            // the four lanes l, l+8, l+16, l+24 hold the same columns (row groups 0..3): warp-shuffle reduction in a fixed order
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                sx[k] += __shfl_xor_sync(0xffffffffu, sx[k], 8);
                sq[k] += __shfl_xor_sync(0xffffffffu, sq[k], 8);
                sx[k] += __shfl_xor_sync(0xffffffffu, sx[k], 16);
                sq[k] += __shfl_xor_sync(0xffffffffu, sq[k], 16);
            }
            if (rsel == 0 && col_ok) {
#pragma unroll
                for (int k = 0; k < 4; ++k) stat_row[col0 + cl + k] = make_double2(sx[k], sq[k]);
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------ the kernel
// PAIR (PRO_F16 operands staged by TMA, MT = 1): the two CTAs of a cluster of 2 form a tcgen05 CTA pair.  CTA rank 0
// issues ONE stream of M = 256 MMAs (.cta_group::2): rows 0-127 are its own time tile, rows 128-255 the peer's, and every
// SM holds -- and fetches from L2 -- only HALF of each weight stage (its NC / 2 output channels; the hardware exchanges
// the halves between the two SMs).  The one-tile kernel is bound by the weight stream (each CTA re-reads the conv's
// whole N-chunk of weights for 128 rows: L2 -> SM traffic and the shared-memory writes of the stages); the pair halves
// both per SM.  Hand-offs: each CTA's TMA / weight copies complete on its own barriers; the peer forwards "my stage has
// landed" to barriers of rank 0 (a_peer / b_peer), rank 0's tcgen05.commit multicasts "stage free" / "accumulator
// complete" to both CTAs.
// threads of a one-tile CTA: 4 MT producer / epilogue warps, MMA warp, loader warp and (MT = 1) four more epilogue warps --
// the epilogue is a third of a CTA's life, and a second warp per tensor-memory lane quarter takes every other 32-column pass
constexpr int conv_threads(int MT) { return 128 * MT + 64 + (MT == 1 ? 128 : 0); }

template <int MODE, int MT, bool PAIR = false>
__global__ void __launch_bounds__(conv_threads(MT)) conv_umma_kernel(const ConvParams p, const __grid_constant__ CUtensorMap a_map,
                                                                     const __grid_constant__ CUtensorMap b_map)
{
    static_assert(!PAIR || (MT == 1 && MODE == PRO_F16), "CTA pairs: one M-tile per CTA, TMA-staged fp16 operand");
    constexpr int N_PRODUCERS = 128 * MT;
    constexpr int MMA_WARP    = 4 * MT;
    constexpr int ROWS_CTA    = TILE_M * MT;
    // warps MMA_WARP + 2 .. + 5, when the launch has them (ConvParams::epi8): epilogue only
    const bool EXTRA_EPI      = MT == 1 && blockDim.x > (unsigned)(N_PRODUCERS + 64);
    const int N_EPI           = N_PRODUCERS + (EXTRA_EPI ? 128 : 0);
    const int EPI_STEP        = EXTRA_EPI ? 2 : 1;
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *bars       = reinterpret_cast<uint64_t *>(smem);
    uint64_t *a_full     = bars;                                  // [MAX_A_STAGES]
    uint64_t *a_empty    = bars + MAX_A_STAGES;                   // [MAX_A_STAGES]
    uint64_t *b_full     = bars + 2 * MAX_A_STAGES;               // [MAX_B_STAGES]
    uint64_t *b_empty    = bars + 2 * MAX_A_STAGES + MAX_B_STAGES;
    uint64_t *acc_full   = bars + 2 * MAX_A_STAGES + 2 * MAX_B_STAGES;
    uint64_t *a_land     = acc_full + 1;                            // [MAX_A_STAGES] TMA mode: the box of a stage has landed
    uint64_t *a_peer     = a_land + MAX_A_STAGES;                   // [MAX_A_STAGES] pair, rank 0: the peer's A stage is ready
    uint64_t *b_peer     = a_peer + MAX_A_STAGES;                   // [MAX_B_STAGES] pair, rank 0: the peer's weight half has landed
    uint32_t *tmem_slot  = reinterpret_cast<uint32_t *>(smem + 336);

    const int tid  = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;

    const int Cin    = p.Cin;
    const int NC     = p.NC;
    const int ntaps  = p.ntaps;
    const int kc_max = Cin < KCHUNK ? Cin : KCHUNK;
    const int nkc_a  = (Cin + KCHUNK - 1) / KCHUNK;                // K-chunks of the conv itself (ntaps taps each)
    const int nkc    = nkc_a + (p.Cin_b + KCHUNK - 1) / KCHUNK;    // + K-chunks of the folded 1-tap source (ConvParams::xb)
    // channels / taps of K-chunk c, and the stage row its first tap starts at (the folded source reads the output row itself)
    auto chunk_kc   = [&](int c) { return c < nkc_a ? min(KCHUNK, Cin - c * KCHUNK) : min(KCHUNK, p.Cin_b - (c - nkc_a) * KCHUNK); };
    auto chunk_taps = [&](int c) { return c < nkc_a ? ntaps : 1; };
    const uint32_t lbo_a         = (uint32_t)p.a_rows * 16u;
    const uint32_t a_stage_bytes = (uint32_t)(kc_max >> 3) * lbo_a;
    const int NB                 = PAIR ? NC / 2 : NC;            // weight rows (output channels) held by this CTA
    const uint32_t lbo_b         = (uint32_t)NB * 16u;
    const uint32_t b_stage_bytes = (uint32_t)kc_max * NB * 2u;
    const uint32_t smem_base     = smem_u32(smem);
    const uint32_t a_base        = smem_base + SMEM_HEADER;
    const uint32_t b_base        = a_base + p.a_stages * a_stage_bytes;
    const bool tma = PAIR || (MODE == PRO_F16 && p.use_tma);

    // ---- thread-block cluster: CL CTAs = CL time tiles of the same N-chunk share every weight stage.  Each CTA fetches
    //      1/CL of a stage and multicasts it to all of them, so the L2 -> SM weight traffic (the bound of these kernels:
    //      a 128-row tile uses every weight byte once) drops by CL.  CTAs that pad the grid recompute the last tile and
    //      store nothing: they must keep consuming the shared stages. ----
    const int CL         = PAIR ? 1 : p.cluster;
    const bool clustered = PAIR || CL > 1;
    const uint32_t crank = clustered ? cluster_ctarank() : 0u;
    const uint16_t cmask = (uint16_t)((1u << CL) - 1u);
    const bool live      = (int)blockIdx.x < p.n_tiles;

    // ---- which tile of which utterance ----
    const int tile    = live ? (int)blockIdx.x : p.n_tiles - 1;
    const int u       = find_segment_warp(p.tile_start, p.B, tile);
    const int t0      = (tile - __ldg(p.tile_start + u)) * ROWS_CTA;
    const int seg_f0  = __ldg(p.seg_start + u);
    const int seg_len = (__ldg(p.seg_start + u + 1) - seg_f0) * p.rate_in;
    const size_t seg_row0 = (size_t)seg_f0 * p.rate_in;
    const int nchunk  = blockIdx.y;
    const TileCoord tc = {u, t0, seg_len, seg_row0, nchunk};

    if (tid == 0) {
        for (int s = 0; s < p.a_stages; ++s) {
            mbar_init(smem_u32(a_full + s), tma ? 1u : (uint32_t)N_PRODUCERS);
            mbar_init(smem_u32(a_empty + s), 1);
            mbar_init(smem_u32(a_land + s), 1);
        }
        if (tma) tma_prefetch_desc(&a_map);
        if (tma && p.Cin_b) tma_prefetch_desc(&b_map);
        for (int s = 0; s < p.b_stages; ++s) {
            mbar_init(smem_u32(b_full + s), 1);
            mbar_init(smem_u32(b_empty + s), (uint32_t)CL);       // a stage is free when every CTA of the cluster has used it
        }
        if (PAIR) {
            for (int s = 0; s < p.a_stages; ++s) mbar_init(smem_u32(a_peer + s), 1);
            for (int s = 0; s < p.b_stages; ++s) mbar_init(smem_u32(b_peer + s), 1);
        }
        mbar_init(smem_u32(acc_full), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) {
        if (PAIR) tmem_alloc2(smem_u32(tmem_slot), (uint32_t)p.tmem_cols);
        else tmem_alloc(smem_u32(tmem_slot), (uint32_t)p.tmem_cols);
    }
    tc_fence_before_sync();
    if (clustered) cluster_sync_all(); else __syncthreads();     // peers' barriers are initialised before anything is multicast
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    const uint32_t acc_stride = (uint32_t)((NC + 31) & ~31);   // tensor-memory columns per M-tile accumulator
    if (warp < MMA_WARP) {
        // =================== A producers ===================
        const int need_rows = ROWS_CTA + (ntaps - 1) * p.tap_step;
        if (tma) {
            // TMA mode: one elected thread asks for the halo tile of every 64-channel chunk as ONE 3-D box
            // (8 channels x a_rows rows x 8 channel groups -> exactly the no-swizzle K-major stage layout); rows beyond
            // the buffer arrive as zeros.  Rows outside THIS utterance (the packed neighbours) are zeroed by warp 1 after
            // the box has landed -- only the first / last tile of an utterance has any -- and warp 1 then hands the
            // stage to the MMA warp.  Warps 2.. wait for the epilogue.
            const int row_first = t0 + p.tap_off0;                          // utterance row of stage row 0
            if (warp == 0) {
                if (lane == 0) {
                    const int gr = (int)((long long)seg_row0 - p.tma_row0) + row_first;
                    const uint32_t box_bytes = 8u * lbo_a;
                    int sa = 0;
                    uint32_t ph = 0;
                    for (int c = 0; c < nkc; ++c) {
                        mbar_wait(smem_u32(a_empty + sa), ph ^ 1u, p.err_flag);
                        mbar_arrive_expect_tx(smem_u32(a_land + sa), box_bytes);
                        if (c < nkc_a) tma_load_3d(a_base + sa * a_stage_bytes, &a_map, 0, gr, c * (KCHUNK / 8), smem_u32(a_land + sa));
                        else tma_load_3d(a_base + sa * a_stage_bytes, &b_map, 0, gr, (c - nkc_a) * (KCHUNK / 8), smem_u32(a_land + sa));
                        if (++sa == p.a_stages) { sa = 0; ph ^= 1u; }
                    }
                }
                __syncwarp();
            } else if (warp == 1) {
                const int lo_end = min(max(-row_first, 0), need_rows);           // stage rows [0, lo_end) precede the utterance
                const int hi_beg = min(max(seg_len - row_first, 0), need_rows);  // stage rows [hi_beg, need_rows) follow it
                const int nbad = lo_end + (need_rows - hi_beg);
                int sa = 0;
                uint32_t ph = 0;
                for (int c = 0; c < nkc; ++c) {
                    mbar_wait(smem_u32(a_land + sa), ph, p.err_flag);
                    if (nbad > 0) {
                        uint8_t *stage = smem + SMEM_HEADER + (size_t)sa * a_stage_bytes;
                        for (int i = lane; i < nbad * 8; i += 32) {
                            const int g = i & 7, k = i >> 3;
                            const int rho = k < lo_end ? k : hi_beg + (k - lo_end);
                            *reinterpret_cast<uint4 *>(stage + (size_t)g * lbo_a + (size_t)rho * 16) = make_uint4(0u, 0u, 0u, 0u);
                        }
                        fence_proxy_async_smem();
                    }
                    __syncwarp();
                    if (lane == 0) {
                        if (PAIR && crank != 0) mbar_arrive_cluster(mapa_u32(smem_u32(a_peer + sa), 0u));   // to rank 0's issuer
                        else mbar_arrive(smem_u32(a_full + sa));
                    }
                    if (++sa == p.a_stages) { sa = 0; ph ^= 1u; }
                }
            }
        } else {
            int sa = 0;
            uint32_t ph = 0;
            for (int c = 0; c < nkc; ++c) {
                mbar_wait(smem_u32(a_empty + sa), ph ^ 1u, p.err_flag);
                produce_chunk<MODE, N_PRODUCERS, 8>(p, tc, c, smem + SMEM_HEADER + (size_t)sa * a_stage_bytes, lbo_a, need_rows, tid);
                if (MODE == PRO_F16) {
                    cp_async_mbar_arrive_noinc(smem_u32(a_full + sa));
                } else {
                    fence_proxy_async_smem();
                    mbar_arrive(smem_u32(a_full + sa));
                }
                if (++sa == p.a_stages) { sa = 0; ph ^= 1u; }
            }
        }

    }
    if (warp < MMA_WARP || (EXTRA_EPI && warp >= MMA_WARP + 2)) {
        // =================== epilogue ===================
        mbar_wait(smem_u32(acc_full), 0u, p.err_flag);
        tc_fence_after_sync();
        // every MMA has completed (acc_full), so the operand stages are dead: their memory is the slab
        const int  ew    = warp < MMA_WARP ? warp : warp - 2;      // epilogue warp index 0 .. 4 MT (+ 4) - 1
        const int  mt    = warp < MMA_WARP ? warp >> 2 : 0;        // M-tile this warp drains
        const int  pass0 = warp < MMA_WARP ? 0 : 1;                // the two warps of a lane quarter alternate the passes
        const uint32_t trow = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)mt * acc_stride;
        // (the stage memory doubles as the transpose slabs: no peer may still be multicasting into it -- every stage this
        //  CTA waited for was the last one its peers sent, and they send nothing after the final K-chunk)
        // statistics partials of the 4 * MT lane quarters sit behind the slabs (also dead stage memory)
        const int N_SLABS = 4 * MT + (EXTRA_EPI ? 4 : 0);
        double2 *stat_all = reinterpret_cast<double2 *>(smem + SMEM_HEADER + (size_t)N_SLABS * SLAB_BYTES);
        constexpr bool STATS = MODE == PRO_F16 || MODE == PRO_NORM;      // the decoder convs (conv_umma_plan enforces it)
        const int quarter_row = warp < MMA_WARP ? warp : (warp & 3);    // both warps of a quarter fill the same row (disjoint columns)
        double2 *stat_row = STATS && p.stats_out ? stat_all + (size_t)quarter_row * NC : nullptr;
        epilogue_tile<STATS>(p, trow, reinterpret_cast<float *>(smem + SMEM_HEADER + (size_t)ew * SLAB_BYTES), lane,
                      t0 + mt * TILE_M + (warp & 3) * 32, live ? seg_len : 0, seg_row0, nchunk, NC, pass0, EPI_STEP, stat_row);
        if (STATS && p.stats_out) {
            // all lane quarters of the tile, summed in a fixed order, one (sum, sum of squares) per output channel
            asm volatile("bar.sync 1, %0;" ::"r"(N_EPI) : "memory");
            const int et = warp < MMA_WARP ? tid : tid - 64;
            if (live)
                for (int c = et; c < NC; c += N_EPI) {
                    double a = 0.0, b = 0.0;
#pragma unroll
                    for (int q = 0; q < 4 * MT; ++q) { const double2 v = stat_all[(size_t)q * NC + c]; a += v.x; b += v.y; }
                    p.stats_out[(size_t)blockIdx.x * p.Cout + (size_t)nchunk * NC + c] = make_double2(a, b);
                }
        }
    } else if (warp == MMA_WARP && PAIR && crank != 0) {
        // =================== pair, rank 1: no MMAs to issue -- tell rank 0 when each weight half has landed ===================
        const int b_stages = p.b_stages;
        int sb = 0;
        uint32_t phb = 0;
        for (int c = 0; c < nkc; ++c)
            for (int a = 0; a < chunk_taps(c); ++a) {
                mbar_wait(smem_u32(b_full + sb), phb, p.err_flag);
                if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(b_peer + sb), 0u));
                __syncwarp();
                if (++sb == b_stages) { sb = 0; phb ^= 1u; }
            }
    } else if (warp == MMA_WARP) {
        // =================== MMA issuer (one elected lane of a converged warp) ===================
        const uint32_t leader = elect_one();
        const uint32_t idesc = PAIR ? make_idesc_mn(2 * TILE_M, NC) : make_idesc(NC);
        // descriptors advance by plain additions on the 14-bit start-address field (no carry out of
        // it: shared memory is < 256 KB); stage indices / phases are counted, not divided
        const uint64_t a_kstep = (uint64_t)((2u * lbo_a) >> 4), b_kstep = (uint64_t)((2u * lbo_b) >> 4);
        const int a_stages = p.a_stages, b_stages = p.b_stages, tap_bytes = p.tap_step * 16;
        int sa = 0, sb = 0;
        uint32_t pha = 0, phb = 0, accum = 0;
        for (int c = 0; c < nkc; ++c) {
            const int ksteps = chunk_kc(c) >> 4;
            const int ctaps = chunk_taps(c);
            mbar_wait(smem_u32(a_full + sa), pha, p.err_flag);
            if (PAIR) mbar_wait_cluster(smem_u32(a_peer + sa), pha, p.err_flag);
            if (MODE == PRO_F16) fence_proxy_async_smem();    // cp.async wrote through the generic proxy
            tc_fence_after_sync();
            // (the folded 1-tap source reads the output row: stage row -tap_off0)
            uint32_t a_tap = a_base + sa * a_stage_bytes + (c < nkc_a ? 0u : (uint32_t)(-p.tap_off0) * 16u);
            for (int a = 0; a < ctaps; ++a, a_tap += tap_bytes) {
                mbar_wait(smem_u32(b_full + sb), phb, p.err_flag);
                if (PAIR) mbar_wait_cluster(smem_u32(b_peer + sb), phb, p.err_flag);
                tc_fence_after_sync();
                if (leader) {
                    const uint64_t bdesc0 = make_smem_desc(b_base + sb * b_stage_bytes, lbo_b, 128u);
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        uint64_t adesc = make_smem_desc(a_tap + (uint32_t)mt * (TILE_M * 16u), lbo_a, 128u);
                        uint64_t bdesc = bdesc0;
                        uint32_t acc = accum;
                        for (int kk = 0; kk < ksteps; ++kk, adesc += a_kstep, bdesc += b_kstep) {
                            if (PAIR) umma2_f16(tmem_base, adesc, bdesc, idesc, acc);
                            else umma_f16(tmem_base + (uint32_t)mt * acc_stride, adesc, bdesc, idesc, acc);
                            acc = 1u;
                        }
                    }
                    if (PAIR) umma2_commit(smem_u32(b_empty + sb));
                    else if (CL > 1) umma_commit_multicast(smem_u32(b_empty + sb), cmask);
                    else umma_commit(smem_u32(b_empty + sb));
                }
                accum = 1u;
                __syncwarp();
                if (++sb == b_stages) { sb = 0; phb ^= 1u; }
            }
            if (leader) {
                if (PAIR) umma2_commit(smem_u32(a_empty + sa));
                else umma_commit(smem_u32(a_empty + sa));
            }
            __syncwarp();
            if (++sa == a_stages) { sa = 0; pha ^= 1u; }
        }
        if (leader) {
            if (PAIR) umma2_commit(smem_u32(acc_full));
            else umma_commit(smem_u32(acc_full));
        }
        __syncwarp();
    } else {
        // =================== weight (B operand) loader ===================
        if (lane == 0) {
            // the packed blocks of one N-chunk are consecutive in (K-chunk, tap) order
            const uint8_t *src = reinterpret_cast<const uint8_t *>(p.w_packed + (size_t)nchunk * (((size_t)Cin * ntaps + (size_t)p.Cin_b) * NC));
            const int b_stages = p.b_stages;
            int sb = 0;
            uint32_t phb = 0;
            for (int c = 0; c < nkc; ++c) {
                const uint32_t bytes = (uint32_t)chunk_kc(c) * NC * 2u;
                const uint32_t part = bytes / (uint32_t)CL;          // this CTA's share of the stage (a multiple of 16 bytes)
#ifdef ZVX_WHATIF_HALF_WEIGHTS
                const uint32_t wbytes = (bytes / 32u) * 16u;         // timing experiment only (wrong results): half the weight stream
#else
                const uint32_t wbytes = bytes;
#endif
                for (int a = 0; a < chunk_taps(c); ++a, src += bytes) {
                    // b_empty counts the commits of ALL CTAs of the cluster: the share goes into every peer's stage
                    mbar_wait(smem_u32(b_empty + sb), phb ^ 1u, p.err_flag);
                    if (PAIR) {
                        // this CTA's half of the block (p.w_packed is the pair layout: [half][group][NC / 2][8]): one copy
                        const uint32_t half_bytes = bytes / 2u;
                        mbar_arrive_expect_tx(smem_u32(b_full + sb), half_bytes);
                        bulk_copy_g2s(b_base + sb * b_stage_bytes, src + crank * half_bytes, half_bytes, smem_u32(b_full + sb));
                        if (++sb == b_stages) { sb = 0; phb ^= 1u; }
                        continue;
                    }
                    mbar_arrive_expect_tx(smem_u32(b_full + sb), CL > 1 ? bytes : wbytes);
                    if (CL > 1)
                        bulk_copy_g2s_multicast(b_base + sb * b_stage_bytes + crank * part, src + crank * part, part, smem_u32(b_full + sb), cmask);
                    else
                        bulk_copy_g2s(b_base + sb * b_stage_bytes, src, wbytes, smem_u32(b_full + sb));
                    if (++sb == b_stages) { sb = 0; phb ^= 1u; }
                }
            }
        }
        __syncwarp();
    }

    tc_fence_before_sync();
    if (clustered) cluster_sync_all(); else __syncthreads();     // no CTA exits while a peer may still signal its barriers
    if (warp == MMA_WARP) {
        tc_fence_after_sync();
        if (PAIR) tmem_dealloc2(tmem_base, (uint32_t)p.tmem_cols);
        else tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

// ------------------------------------------------------------------ persistent variant
// Same math, different schedule: one CTA per SM loops over work items (tile, N-chunk) and the
// roles are fully separated -- warps 0-3 produce A stages, warps 4-7 drain accumulators, warp 8
// issues MMAs, warp 9 streams weights.  Two accumulators in tensor memory alternate, so the
// epilogue of item i (HBM traffic: residual read, fp32 write) runs under the main loop of item
// i+1, and with the whole shared memory for one CTA the weight ring is deep enough to cover the
// L2 round trip of cp.async.bulk.  (Two co-resident one-tile CTAs run in lockstep and reach their
// epilogues together; the per-launch timings in profiles/ show conv2-type launches paying their
// whole epilogue.)
constexpr int PK_THREADS   = 448;     // warps 0-3 A producers, 4-7 and 10-13 epilogue, 8 MMA issuer, 9 weight loader
constexpr int PK_EPI_WARPS = 8;       // two warps per tensor-memory lane quarter, alternating 32-column passes

template <int MODE>
__global__ void __launch_bounds__(PK_THREADS, 1) conv_umma_pk_kernel(const ConvParams p, const int total_tiles, const int nchunks)
{
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *bars       = reinterpret_cast<uint64_t *>(smem);
    uint64_t *a_full     = bars;                                  // [MAX_A_STAGES]
    uint64_t *a_empty    = bars + MAX_A_STAGES;                   // [MAX_A_STAGES]
    uint64_t *b_full     = bars + 2 * MAX_A_STAGES;               // [MAX_B_STAGES]
    uint64_t *b_empty    = bars + 2 * MAX_A_STAGES + MAX_B_STAGES;
    uint64_t *acc_full   = bars + 2 * MAX_A_STAGES + 2 * MAX_B_STAGES;        // [2]
    uint64_t *acc_empty  = acc_full + 2;                                      // [2]
    uint32_t *tmem_slot  = reinterpret_cast<uint32_t *>(smem + 248);

    const int tid  = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;

    const int Cin    = p.Cin;
    const int NC     = p.NC;
    const int ntaps  = p.ntaps;
    const int kc_max = Cin < KCHUNK ? Cin : KCHUNK;
    const int nkc    = (Cin + KCHUNK - 1) / KCHUNK;
    const uint32_t lbo_a         = (uint32_t)p.a_rows * 16u;
    const uint32_t a_stage_bytes = (uint32_t)(kc_max >> 3) * lbo_a;
    const uint32_t lbo_b         = (uint32_t)NC * 16u;
    const uint32_t b_stage_bytes = (uint32_t)kc_max * NC * 2u;
    const uint32_t smem_base     = smem_u32(smem);
    const uint32_t a_base        = smem_base + SMEM_HEADER;
    const uint32_t b_base        = a_base + p.a_stages * a_stage_bytes;
    const int total_items        = total_tiles * nchunks;
    const uint32_t acc_stride    = (uint32_t)((NC + 31) & ~31);

    auto coord = [&](int w) {
        TileCoord tc;
        const int tile = w / nchunks;
        tc.nchunk  = w - tile * nchunks;
        tc.u       = find_segment_warp(p.tile_start, p.B, tile);
        tc.t0      = (tile - __ldg(p.tile_start + tc.u)) * TILE_M;
        const int seg_f0 = __ldg(p.seg_start + tc.u);
        tc.seg_len  = (__ldg(p.seg_start + tc.u + 1) - seg_f0) * p.rate_in;
        tc.seg_row0 = (size_t)seg_f0 * p.rate_in;
        return tc;
    };

    if (tid == 0) {
        for (int s = 0; s < p.a_stages; ++s) {
            mbar_init(smem_u32(a_full + s), 128);
            mbar_init(smem_u32(a_empty + s), 1);
        }
        for (int s = 0; s < p.b_stages; ++s) {
            mbar_init(smem_u32(b_full + s), 1);
            mbar_init(smem_u32(b_empty + s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(acc_full + s), 1);
            mbar_init(smem_u32(acc_empty + s), 32 * PK_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) tmem_alloc(smem_u32(tmem_slot), (uint32_t)p.tmem_cols);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        // =================== A producers ===================
        const int need_rows = TILE_M + (ntaps - 1) * p.tap_step;
        int sa = 0;
        uint32_t ph = 0;
        for (int w = blockIdx.x; w < total_items; w += gridDim.x) {
            const TileCoord tc = coord(w);
            for (int c = 0; c < nkc; ++c) {
                mbar_wait(smem_u32(a_empty + sa), ph ^ 1u, p.err_flag);
                produce_chunk<MODE, 128, 8>(p, tc, c, smem + SMEM_HEADER + (size_t)sa * a_stage_bytes, lbo_a, need_rows, tid);
                if (MODE == PRO_F16) {
                    cp_async_mbar_arrive_noinc(smem_u32(a_full + sa));
                } else {
                    fence_proxy_async_smem();
                    mbar_arrive(smem_u32(a_full + sa));
                }
                if (++sa == p.a_stages) { sa = 0; ph ^= 1u; }
            }
        }
    } else if (warp < 8 || warp >= 10) {
        // =================== epilogue ===================
        // (runs under the next item's main loop, so its transpose slabs sit behind the operand stages).  A warp reads
        // the tensor-memory lanes of quarter warp % 4; the two warps of a quarter alternate the 32-column passes.
        const int eidx = warp < 8 ? warp - 4 : warp - 10 + 4;      // 0..7
        const int egrp = eidx >> 2;                                // which of the two warps of the lane quarter
        float *slab = reinterpret_cast<float *>(smem + SMEM_HEADER + (size_t)p.a_stages * a_stage_bytes + (size_t)p.b_stages * b_stage_bytes +
                                                (size_t)eidx * SLAB_BYTES);
        int n = 0;
        for (int w = blockIdx.x; w < total_items; w += gridDim.x, ++n) {
            const TileCoord tc = coord(w);
            const int buf = n & 1;
            mbar_wait(smem_u32(acc_full + buf), (uint32_t)(n >> 1) & 1u, p.err_flag);
            tc_fence_after_sync();
            const uint32_t trow = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)buf * acc_stride;
            epilogue_tile<false>(p, trow, slab, lane, tc.t0 + (warp & 3) * 32, tc.seg_len, tc.seg_row0, tc.nchunk, NC, egrp, 2);
            tc_fence_before_sync();
            mbar_arrive(smem_u32(acc_empty + buf));
        }
    } else if (warp == 8) {
        // =================== MMA issuer ===================
        const uint32_t leader = elect_one();
        const uint32_t idesc = make_idesc(NC);
        const uint64_t a_kstep = (uint64_t)((2u * lbo_a) >> 4), b_kstep = (uint64_t)((2u * lbo_b) >> 4);
        const int a_stages = p.a_stages, b_stages = p.b_stages, tap_bytes = p.tap_step * 16;
        int sa = 0, sb = 0, n = 0;
        uint32_t pha = 0, phb = 0;
        for (int w = blockIdx.x; w < total_items; w += gridDim.x, ++n) {
            const int buf = n & 1;
            mbar_wait(smem_u32(acc_empty + buf), ((uint32_t)(n >> 1) & 1u) ^ 1u, p.err_flag);
            tc_fence_after_sync();
            const uint32_t dcol = tmem_base + (uint32_t)buf * acc_stride;
            uint32_t accum = 0;
            for (int c = 0; c < nkc; ++c) {
                const int ksteps = min(KCHUNK, Cin - c * KCHUNK) >> 4;
                mbar_wait(smem_u32(a_full + sa), pha, p.err_flag);
                if (MODE == PRO_F16) fence_proxy_async_smem();    // cp.async wrote through the generic proxy
                tc_fence_after_sync();
                uint32_t a_tap = a_base + sa * a_stage_bytes;
                for (int a = 0; a < ntaps; ++a, a_tap += tap_bytes) {
                    mbar_wait(smem_u32(b_full + sb), phb, p.err_flag);
                    tc_fence_after_sync();
                    if (leader) {
                        uint64_t adesc = make_smem_desc(a_tap, lbo_a, 128u);
                        uint64_t bdesc = make_smem_desc(b_base + sb * b_stage_bytes, lbo_b, 128u);
                        for (int kk = 0; kk < ksteps; ++kk, adesc += a_kstep, bdesc += b_kstep) {
                            umma_f16(dcol, adesc, bdesc, idesc, accum);
                            accum = 1u;
                        }
                        umma_commit(smem_u32(b_empty + sb));
                    }
                    accum = 1u;
                    __syncwarp();
                    if (++sb == b_stages) { sb = 0; phb ^= 1u; }
                }
                if (leader) umma_commit(smem_u32(a_empty + sa));
                __syncwarp();
                if (++sa == a_stages) { sa = 0; pha ^= 1u; }
            }
            if (leader) umma_commit(smem_u32(acc_full + buf));
            __syncwarp();
        }
    } else {
        // =================== weight (B operand) loader ===================
        if (lane == 0) {
            const int b_stages = p.b_stages;
            int sb = 0;
            uint32_t phb = 0;
            for (int w = blockIdx.x; w < total_items; w += gridDim.x) {
                const int nchunk = w % nchunks;
                const uint8_t *src = reinterpret_cast<const uint8_t *>(p.w_packed + (size_t)nchunk * ((size_t)Cin * ntaps * NC));
                for (int c = 0; c < nkc; ++c) {
                    const uint32_t bytes = (uint32_t)min(KCHUNK, Cin - c * KCHUNK) * NC * 2u;
                    for (int a = 0; a < ntaps; ++a, src += bytes) {
                        mbar_wait(smem_u32(b_empty + sb), phb ^ 1u, p.err_flag);
                        mbar_arrive_expect_tx(smem_u32(b_full + sb), bytes);
                        bulk_copy_g2s(b_base + sb * b_stage_bytes, src, bytes, smem_u32(b_full + sb));
                        if (++sb == b_stages) { sb = 0; phb ^= 1u; }
                    }
                }
            }
        }
        __syncwarp();
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

// ------------------------------------------------------------------ host side
static int round_a_rows(int need_rows, int kc_max)
{
    // rows per A stage, padded so that the 16-byte stores of one quarter-warp hit 32 distinct
    // banks: with G = kc/8 groups, (rows*j + r) mod 8 must be distinct for j < G, r < 8/G.
    const int G = kc_max >> 3;
    const int want = (G >= 8) ? 1 : (G == 4) ? 2 : 4;
    int r = need_rows;
    while ((r & 7) != want) ++r;
    return r;
}

size_t conv_umma_plan(ConvParams &p, size_t smem_budget)
{
    const int kc_max    = p.Cin < KCHUNK ? p.Cin : KCHUNK;
    const int need_rows = TILE_M * p.mt + (p.ntaps - 1) * p.tap_step;
    // TMA boxes: at most 256 rows, always 8 channel groups (a partial last chunk is zero-filled), rows exactly the box
    if (p.stats_out && p.pro_mode != PRO_F16 && p.pro_mode != PRO_NORM) p.stats_out = nullptr;
    if (p.use_tma && (p.pro_mode != PRO_F16 || need_rows > 256 || p.Cin < KCHUNK || p.ldx % 8 || p.x_ch_off % 8)) p.use_tma = 0;
    if (p.pair && (!p.use_tma || p.mt != 1 || p.NC % 16 || p.NC < 32)) p.pair = 0;     // pairs: TMA-staged operand, N a multiple of 16
    p.a_rows            = p.use_tma ? need_rows : round_a_rows(need_rows, kc_max);
    const size_t a_stage = (size_t)(kc_max >> 3) * p.a_rows * 16;
    const size_t b_stage = (size_t)kc_max * p.NC * 2 / (p.pair ? 2 : 1);
    if (p.Cin_b && (!p.use_tma || p.Cin_b < KCHUNK || p.Cin_b % 8 || p.ldxb % 8)) p.Cin_b = -1;      // cannot fold: the caller must check
    const int nkc_b      = p.Cin_b > 0 ? (p.Cin_b + KCHUNK - 1) / KCHUNK : 0;
    const int nkc        = (p.Cin + KCHUNK - 1) / KCHUNK + nkc_b;
    const int nb_total   = (nkc - nkc_b) * p.ntaps + nkc_b;
    int as = nkc < 2 ? 1 : 2;
    int bs = nb_total < 2 ? 1 : 2;
    // grow B first (weights are the longer stream), then A, while it fits
    while (true) {
        bool grew = false;
        if (bs < MAX_B_STAGES && bs < nb_total && SMEM_HEADER + as * a_stage + (bs + 1) * b_stage <= smem_budget) {
            ++bs;
            grew = true;
        }
        if (as < MAX_A_STAGES && as < nkc && as < 3 && SMEM_HEADER + (as + 1) * a_stage + bs * b_stage <= smem_budget) {
            ++as;
            grew = true;
        }
        if (!grew) break;
    }
    p.a_stages = as;
    p.b_stages = bs;
    int cols = 32;
    while (cols < p.mt * ((p.NC + 31) & ~31)) cols <<= 1;
    p.tmem_cols = cols;
    // the epilogue reuses the stage memory for its transpose slabs (one per producer warp)
    const size_t stages = as * a_stage + bs * b_stage,
                 slabs = (size_t)(4 * p.mt + (p.mt == 1 && p.epi8 ? 4 : 0)) * SLAB_BYTES + (p.stats_out ? (size_t)4 * p.mt * p.NC * sizeof(double2) : 0);
    return SMEM_HEADER + (stages > slabs ? stages : slabs);
}

// smem plan of the persistent kernel: the whole SM for one CTA, weight ring as deep as it fits
size_t conv_umma_pk_plan(ConvParams &p, size_t smem_budget)
{
    p.mt = 1;
    const int kc_max    = p.Cin < KCHUNK ? p.Cin : KCHUNK;
    const int need_rows = TILE_M + (p.ntaps - 1) * p.tap_step;
    p.a_rows            = round_a_rows(need_rows, kc_max);
    const size_t a_stage = (size_t)(kc_max >> 3) * p.a_rows * 16;
    const size_t b_stage = (size_t)kc_max * p.NC * 2;
    int as = 3, bs = 2;
    smem_budget -= PK_EPI_WARPS * SLAB_BYTES;      // epilogue transpose slabs behind the stages
    while (bs < MAX_B_STAGES && SMEM_HEADER + as * a_stage + (bs + 1) * b_stage <= smem_budget) ++bs;
    while (as < MAX_A_STAGES && SMEM_HEADER + (as + 1) * a_stage + bs * b_stage <= smem_budget) ++as;
    p.a_stages = as;
    p.b_stages = bs;
    int cols = 32;
    while (cols < 2 * ((p.NC + 31) & ~31)) cols <<= 1;
    p.tmem_cols = cols;
    return SMEM_HEADER + as * a_stage + bs * b_stage + PK_EPI_WARPS * SLAB_BYTES;
}

template <int MODE>
static cudaError_t launch_pk(const ConvParams &p, int total_tiles, int num_sms, size_t smem, cudaStream_t st)
{
    const int nchunks = p.Cout / p.NC;
    const int items = total_tiles * nchunks;
    const int grid = items < num_sms ? items : num_sms;
    conv_umma_pk_kernel<MODE><<<grid, PK_THREADS, smem, st>>>(p, total_tiles, nchunks);
    return cudaGetLastError();
}

// total_tiles counts 128-row tiles
cudaError_t conv_umma_pk_launch(const ConvParams &p, int total_tiles, int num_sms, size_t smem, cudaStream_t st)
{
    switch (p.pro_mode) {
        case PRO_F16:   return launch_pk<PRO_F16>(p, total_tiles, num_sms, smem, st);
        case PRO_CVT:   return launch_pk<PRO_CVT>(p, total_tiles, num_sms, smem, st);
        case PRO_LRELU: return launch_pk<PRO_LRELU>(p, total_tiles, num_sms, smem, st);
        case PRO_NORM:  return launch_pk<PRO_NORM>(p, total_tiles, num_sms, smem, st);
        case PRO_MEL:   return launch_pk<PRO_MEL>(p, total_tiles, num_sms, smem, st);
        case PRO_SUM3:  return launch_pk<PRO_SUM3>(p, total_tiles, num_sms, smem, st);
        case PRO_SUM3H: return launch_pk<PRO_SUM3H>(p, total_tiles, num_sms, smem, st);
    }
    return cudaErrorInvalidValue;
}

// launch with a thread-block cluster of p.cluster CTAs along x (1: plain launch)
template <typename K>
static cudaError_t launch_clustered(K kernel, dim3 grid, int threads, size_t smem, cudaStream_t st, const ConvParams &p, const CUtensorMap &a_map,
                                    const CUtensorMap &b_map)
{
    if (p.cluster <= 1) {
        kernel<<<grid, threads, smem, st>>>(p, a_map, b_map);
        return cudaGetLastError();
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(threads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)p.cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, p, a_map, b_map);
}

// Tensor map of a PRO_F16 operand buffer for the TMA-staged A tiles: the [rows][ldx] fp16 matrix (first channel x_ch_off)
// seen as (8 channels, rows, C/8 channel groups) with strides (2, 2 ldx, 16) bytes, box (8, box_rows, 8), no swizzle:
// a box lands as [group][row][8 channels] = the K-major no-swizzle operand layout with LBO = box_rows * 16 bytes.
static cudaError_t make_a_map(const ConvParams &p, long long rows_total, CUtensorMap *out, bool second = false)
{
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        const cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) return cudaErrorNotSupported;
        encode = reinterpret_cast<encode_fn>(fn);
    }
    // (second: the folded 1-tap source ConvParams::xb, same rows, same box)
    const int ld = second ? p.ldxb : p.ldx, cin = second ? p.Cin_b : p.Cin;
    void *base = second ? const_cast<__half *>(reinterpret_cast<const __half *>(p.xb) + (size_t)p.tma_row0 * ld)
                        : const_cast<__half *>(reinterpret_cast<const __half *>(p.x) + (size_t)p.tma_row0 * ld + p.x_ch_off);
    const cuuint64_t dims[3] = {8, (cuuint64_t)rows_total, (cuuint64_t)(cin / 8)};
    const cuuint64_t strides[2] = {(cuuint64_t)ld * 2, 16};
    const cuuint32_t box[3] = {8, (cuuint32_t)p.a_rows, 8};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

template <int MODE, int MT>
static cudaError_t launch_mode(const ConvParams &p, int total_tiles, size_t smem, cudaStream_t st)
{
    ConvParams q = p;
    if (q.cluster != 2 && q.cluster != 4) q.cluster = 1;
    if (q.pair) q.cluster = 2;
    q.n_tiles = total_tiles;
    dim3 grid((total_tiles + q.cluster - 1) / q.cluster * q.cluster, p.Cout / p.NC, 1);
    CUtensorMap a_map, b_map;
    memset(&a_map, 0, sizeof a_map);
    memset(&b_map, 0, sizeof b_map);
    if (MODE == PRO_F16 && q.use_tma) {
        cudaError_t e = make_a_map(q, q.tma_rows, &a_map);
        if (e != cudaSuccess) return e;
        if (q.Cin_b > 0 && (e = make_a_map(q, q.tma_rows, &b_map, true)) != cudaSuccess) return e;
    } else {
        if (q.Cin_b > 0) return cudaErrorInvalidValue;       // the folded source exists in TMA mode only (the caller checks conv_umma_plan)
        q.use_tma = 0;
    }
    if constexpr (MODE == PRO_F16 && MT == 1) {
        if (q.pair) return launch_clustered(conv_umma_kernel<MODE, MT, true>, grid, q.epi8 ? conv_threads(MT) : 128 * MT + 64, smem, st, q, a_map, b_map);
    }
    return launch_clustered(conv_umma_kernel<MODE, MT>, grid, q.epi8 && MT == 1 ? conv_threads(MT) : 128 * MT + 64, smem, st, q, a_map, b_map);
}

template <int MODE>
static cudaError_t init_mode()
{
    cudaError_t e;
    const int kMax = 227 * 1024;
    if ((e = cudaFuncSetAttribute(conv_umma_pk_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMax)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(conv_umma_kernel<MODE, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMax)) != cudaSuccess) return e;
    if constexpr (MODE == PRO_F16) {
        if ((e = cudaFuncSetAttribute(conv_umma_kernel<MODE, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMax)) != cudaSuccess) return e;
    }
    return cudaFuncSetAttribute(conv_umma_kernel<MODE, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMax);
}

cudaError_t conv_umma_init()
{
    cudaError_t e;
    if ((e = init_mode<PRO_F16>()) != cudaSuccess) return e;
    if ((e = init_mode<PRO_CVT>()) != cudaSuccess) return e;
    if ((e = init_mode<PRO_LRELU>()) != cudaSuccess) return e;
    if ((e = init_mode<PRO_NORM>()) != cudaSuccess) return e;
    if ((e = init_mode<PRO_MEL>()) != cudaSuccess) return e;
    if ((e = init_mode<PRO_SUM3>()) != cudaSuccess) return e;
    if ((e = init_mode<PRO_SUM3H>()) != cudaSuccess) return e;
    return cudaSuccess;
}

template <int MODE>
static cudaError_t launch_mt(const ConvParams &p, int total_tiles, size_t smem, cudaStream_t st)
{
    return p.mt == 2 ? launch_mode<MODE, 2>(p, total_tiles, smem, st) : launch_mode<MODE, 1>(p, total_tiles, smem, st);
}

// total_tiles counts tiles of 128 * p.mt rows (p.tile_start must be the matching prefix table)
cudaError_t conv_umma_launch(const ConvParams &p, int total_tiles, size_t smem, cudaStream_t st)
{
    if (p.mt != 1 && p.mt != 2) return cudaErrorInvalidValue;
    switch (p.pro_mode) {
        case PRO_F16:   return launch_mt<PRO_F16>(p, total_tiles, smem, st);
        case PRO_CVT:   return launch_mt<PRO_CVT>(p, total_tiles, smem, st);
        case PRO_LRELU: return launch_mt<PRO_LRELU>(p, total_tiles, smem, st);
        case PRO_NORM:  return launch_mt<PRO_NORM>(p, total_tiles, smem, st);
        case PRO_MEL:   return launch_mt<PRO_MEL>(p, total_tiles, smem, st);
        case PRO_SUM3:  return launch_mt<PRO_SUM3>(p, total_tiles, smem, st);
        case PRO_SUM3H: return launch_mt<PRO_SUM3H>(p, total_tiles, smem, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace zvx
