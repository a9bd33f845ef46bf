// mrf_fused.cu -- one HiFi-GAN MRF residual block (or a chain of its conv pairs) per launch,
// kept on chip for a whole time window.  Replaces HiFiGANResidualBlock
// (/root/reference/src/hifigan.cpp:74-185): 2 x (#pairs) ggml_conv_1d with leaky-ReLU, bias and
// the residual add in between, plus the branch sum / average of hifigan.cpp:300-315 in the
// final epilogue.  Geometry and the reasons for the swapped (weights = A operand) orientation
// are in mrf_fused.cuh.
//
// One CTA (320 threads, 1 per SM) owns one window of WP time steps of one utterance:
//   warps 0-7  prologue + epilogues.  Thread (quarter q = warp & 3, lane) owns accumulator row
//              m = 32 q + lane = (shift s, output channel oc); warps 0-3 take columns 0-127,
//              warps 4-7 columns 128-255.
//   warp  8    MMA issuer (one elected lane): per layer (k + S - 1) * CH/16 tcgen05.mma of
//              shape M=128 (weights window) x N=256 (positions) x K=16.
//   warp  9    weight loader: cp.async.bulk of per-(layer, K-step) chunks into a ring.
// Tensor memory: columns [0,256) = H accumulator (conv1 output), [256,512) = y.  The residual
// stream y stays in tensor memory in fp32 for the whole chain: conv2's MMAs accumulate
// directly on top of it, its bias is added when y is read (cumulative bias, host-prepared).
// Shared memory: two activation buffers (fp16, layouts per mrf_fused.cuh) that alternate as
// MMA B operand / epilogue destination, and the weight ring.
#include "mrf_fused.cuh"
#include "ptx_sm100.cuh"
#include "zvx_common.cuh"
#include "zvx_internal.h"

namespace zvx {

namespace {

constexpr int F_EPI_WARPS = 16;                 // 4 per lane quarter, 64 columns each
constexpr int F_EPI       = F_EPI_WARPS * 32;
constexpr int F_THREADS   = F_EPI + 64;
constexpr int F_COLS_PER_WARP = mrf::NCOL / (F_EPI_WARPS / 4);
constexpr int F_BATCHES   = F_COLS_PER_WARP / 32;
constexpr int F_HEADER    = 512;
constexpr int F_MAX_SLOTS = 8;

template <int CH>
__host__ __device__ constexpr int f_tables_bytes() { return 2 * mrf::Geo<CH>::S * mrf::NCOL * 4; }

__device__ __forceinline__ float lrelu_max(float x, float a)
{
    // max(x, a x) == ggml's max(x,0) + a min(x,0) for 0 < a < 1 (up to the sign of zero)
    return fmaxf(x, __fmul_rn(a, x));
}
__device__ __forceinline__ void epi_bar_sync()
{
    asm volatile("bar.sync 1, %0;" ::"n"(F_EPI) : "memory");
}

template <int CH>
__global__ void __launch_bounds__(F_THREADS, 1) mrf_fused_kernel(const mrf::Params p, const int nslots, const uint32_t slot_bytes)
{
    using G = mrf::Geo<CH>;
    constexpr int S = G::S;
    constexpr int TBL_WORDS = S * mrf::NCOL;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t *bars      = reinterpret_cast<uint64_t *>(smem);
    uint64_t *w_full    = bars;                        // [F_MAX_SLOTS]
    uint64_t *w_empty   = bars + F_MAX_SLOTS;          // [F_MAX_SLOTS]
    uint64_t *acc_full  = bars + 2 * F_MAX_SLOTS;
    uint64_t *act_ready = bars + 2 * F_MAX_SLOTS + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + 256);
    uint32_t *tbl_s     = reinterpret_cast<uint32_t *>(smem + F_HEADER);     // [2][S][NCOL]
    constexpr uint32_t OFF_BUF0 = F_HEADER + f_tables_bytes<CH>();
    constexpr uint32_t OFF_BUF1 = OFF_BUF0 + G::BUF;
    constexpr uint32_t OFF_RING = OFF_BUF1 + G::BUF;

    const int tid  = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;

    const uint32_t smem_base = smem_u32(smem);
    const uint32_t buf0      = smem_base + OFF_BUF0;
    const uint32_t buf1      = smem_base + OFF_BUF1;
    const uint32_t ring      = smem_base + OFF_RING;

    // ---- which window of which utterance ----
    const int win = blockIdx.x;
    const int u   = find_segment(p.win_start, p.B, win);
    const int wi  = win - __ldg(p.win_start + u);
    const int f0  = __ldg(p.seg_start + u);
    const int T   = (__ldg(p.seg_start + u + 1) - f0) * p.rate;
    const size_t row0 = (size_t)f0 * p.rate;
    const int tw  = wi * p.valid - p.halo;             // time of window position 0 (may be < 0)
    const bool interior = tw >= 0 && tw + G::WP <= T;

    // ---- one-time setup: zero both activation buffers (guard rows and never-written rows must
    //      read as finite zeros), barriers, tensor memory ----
    {
        uint4 *z = reinterpret_cast<uint4 *>(smem + OFF_BUF0);
        for (int i = tid; i < 2 * G::BUF / 16; i += F_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (tid == 0) {
        for (int s = 0; s < nslots; ++s) {
            mbar_init(smem_u32(w_full + s), 1);
            mbar_init(smem_u32(w_empty + s), 1);
        }
        mbar_init(smem_u32(acc_full), 1);
        mbar_init(smem_u32(act_ready), F_EPI);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == F_EPI_WARPS) tmem_alloc(smem_u32(tmem_slot), 512u);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < F_EPI_WARPS) {
        // =================== prologue + epilogues ===================
        const int quarter = warp & 3;
        const int part    = warp >> 2;
        const int m       = quarter * 32 + lane;
        const int s       = m / CH;
        const int oc      = m % CH;
        const uint32_t toff = (uint32_t)(oc >> 3) * mrf::LBO_B + (uint32_t)(oc & 7) * 2u;
        const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16);

        // ---- prologue: y window -> tensor memory (fp32), lrelu(y) -> buffer 0 (fp16) ----
        {
            for (int i = tid; i < TBL_WORDS; i += F_EPI) tbl_s[TBL_WORDS + i] = __ldg(p.tbl0 + i);
            epi_bar_sync();
            const float *yin = p.y_in + row0 * CH + oc;
            const uint32_t *tb = tbl_s + TBL_WORDS + s * mrf::NCOL;
            uint8_t *dst = smem + OFF_BUF0 + toff;
#pragma unroll 1
            for (int b = 0; b < F_BATCHES; ++b) {
                const int col0 = part * F_COLS_PER_WARP + b * 32;
                uint32_t v[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int tau = S * (col0 + i) + s;
                    const int t   = tw + tau;
                    float y = 0.f;
                    if (tau < G::WP && t >= 0 && t < T) y = __ldg(yin + (size_t)t * CH);
                    v[i] = __float_as_uint(y);
                }
                tmem_st32(tlane + 256u + (uint32_t)col0, v);
#pragma unroll
                for (int i4 = 0; i4 < 8; ++i4) {
                    const uint4 e4 = *reinterpret_cast<const uint4 *>(tb + col0 + 4 * i4);
                    const uint32_t e[4] = {e4.x, e4.y, e4.z, e4.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const __half h = __float2half_rn(lrelu_max(__uint_as_float(v[4 * i4 + q]), p.in_slope));
                        if (e[q] & mrf::TBL_VALID) *reinterpret_cast<__half *>(dst + mrf::tbl_byte(e[q])) = h;
                    }
                }
            }
            tmem_wait_st();
            fence_proxy_async_smem();
            tc_fence_before_sync();
            mbar_arrive(smem_u32(act_ready));
        }

        // ---- per-layer epilogues ----
#pragma unroll 1
        for (int l = 0; l < p.nlayers; ++l) {
            const mrf::Layer &L = p.L[l];
            const bool last = l == p.nlayers - 1;
            const float bias = __ldg(L.bias + oc);
            const uint32_t acc = tlane + (L.accumulate ? 256u : 0u);
            if (!last) {
                // stage this layer's scatter table while the MMAs run (double-buffered: a warp can only
                // be one layer ahead of the slowest one, which reads the other copy)
                uint32_t *tdst = tbl_s + (l & 1) * TBL_WORDS;
                for (int i = tid; i < TBL_WORDS; i += F_EPI) tdst[i] = __ldg(L.tbl + i);
                epi_bar_sync();
            }
            mbar_wait(smem_u32(acc_full), (uint32_t)l & 1u, p.err_flag);
            tc_fence_after_sync();
            if (!last) {
                const uint32_t *tb = tbl_s + (l & 1) * TBL_WORDS + s * mrf::NCOL;
                uint8_t *dst = smem + ((l & 1) ? OFF_BUF0 : OFF_BUF1) + toff;
                const float slope = L.out_slope;
#pragma unroll 1
                for (int b = 0; b < F_BATCHES; ++b) {
                    const int col0 = part * F_COLS_PER_WARP + b * 32;
                    uint32_t r[32];
                    tmem_ld32(acc + (uint32_t)col0, r);
#pragma unroll
                    for (int i4 = 0; i4 < 8; ++i4) {
                        const uint4 e4 = *reinterpret_cast<const uint4 *>(tb + col0 + 4 * i4);
                        const uint32_t e[4] = {e4.x, e4.y, e4.z, e4.w};
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            float v = lrelu_max(__fadd_rn(__uint_as_float(r[4 * i4 + q]), bias), slope);
                            if (!interior) {
                                const int t = tw + mrf::tbl_tau(e[q]);
                                if (t < 0 || t >= T) v = 0.f;
                            }
                            if (e[q] & mrf::TBL_VALID) *reinterpret_cast<__half *>(dst + mrf::tbl_byte(e[q])) = __float2half_rn(v);
                        }
                    }
                }
                fence_proxy_async_smem();
                tc_fence_before_sync();
                mbar_arrive(smem_u32(act_ready));
            } else {
                // final: y (+ running branch sum) (* 1/num_blocks) -> global fp32
                float *out = p.out + row0 * CH + oc;
                const float *ain = p.acc_in ? p.acc_in + row0 * CH + oc : nullptr;
#pragma unroll 1
                for (int b = 0; b < F_BATCHES; ++b) {
                    const int col0 = part * F_COLS_PER_WARP + b * 32;
                    uint32_t r[32];
                    tmem_ld32(acc + (uint32_t)col0, r);
                    float a[32];
                    if (ain) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const int tau = S * (col0 + i) + s;
                            const int t   = tw + tau;
                            a[i] = (tau >= p.halo && tau < p.halo + p.valid && t < T) ? ain[(size_t)t * CH] : 0.f;
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int tau = S * (col0 + i) + s;
                        const int t   = tw + tau;
                        if (tau >= p.halo && tau < p.halo + p.valid && t < T) {
                            float v = __fadd_rn(__uint_as_float(r[i]), bias);
                            if (ain) v = __fadd_rn(a[i], v);
                            if (p.has_scale) v = __fmul_rn(v, p.scale);
                            out[(size_t)t * CH] = v;
                        }
                    }
                }
            }
        }
    } else if (warp == F_EPI_WARPS) {
        // =================== MMA issuer ===================
        const uint32_t leader = elect_one();
        const uint32_t idesc  = make_idesc_mn(128, mrf::NCOL);
        int it = 0;
#pragma unroll 1
        for (int l = 0; l < p.nlayers; ++l) {
            const mrf::Layer &L = p.L[l];
            const int k  = L.k;
            const int nj = k + S - 1;
            const uint32_t lbo_a = (uint32_t)mrf::tap_blocks(k, S) * CH * 16u;
            const uint32_t ibuf  = (l & 1) ? buf1 : buf0;
            const uint32_t dcol  = tmem_base + (L.accumulate ? 256u : 0u);
            mbar_wait(smem_u32(act_ready), (uint32_t)l & 1u, p.err_flag);
            tc_fence_after_sync();
#pragma unroll 1
            for (int c = 0; c < G::KSTEPS; ++c, ++it) {
                const int slot = it % nslots;
                const uint32_t ph = (uint32_t)(it / nslots) & 1u;
                mbar_wait(smem_u32(w_full + slot), ph, p.err_flag);
                tc_fence_after_sync();
                if (leader) {
                    const uint32_t a_slot = ring + (uint32_t)slot * slot_bytes;
                    const uint32_t b_c    = ibuf + (uint32_t)(2 * c) * mrf::LBO_B + (uint32_t)mrf::GUARD * 16u;
#pragma unroll 1
                    for (int j = 0; j < nj; ++j) {
                        int q, ro;
                        mrf::b_step(k, S, j, q, ro);
                        const uint64_t adesc = make_smem_desc(a_slot + (uint32_t)mrf::a_block(k, S, j) * (CH * 16u), lbo_a, 128u);
                        const uint64_t bdesc = make_smem_desc(b_c + (uint32_t)q * G::SUB + (uint32_t)(ro * 16), mrf::LBO_B, 128u);
                        umma_f16(dcol, adesc, bdesc, idesc, (L.accumulate || c > 0 || j > 0) ? 1u : 0u);
                    }
                    umma_commit(smem_u32(w_empty + slot));
                }
                __syncwarp();
            }
            if (leader) umma_commit(smem_u32(acc_full));
            __syncwarp();
        }
    } else {
        // =================== weight loader ===================
        if (lane == 0) {
            int it = 0;
            for (int l = 0; l < p.nlayers; ++l) {
                const mrf::Layer &L = p.L[l];
                const uint32_t bytes = mrf::chunk_bytes(L.k, S, CH);
                for (int c = 0; c < G::KSTEPS; ++c, ++it) {
                    const int slot = it % nslots;
                    const uint32_t ph = (uint32_t)(it / nslots) & 1u;
                    mbar_wait(smem_u32(w_empty + slot), ph ^ 1u, p.err_flag);
                    mbar_arrive_expect_tx(smem_u32(w_full + slot), bytes);
                    bulk_copy_g2s(ring + (uint32_t)slot * slot_bytes, reinterpret_cast<const uint8_t *>(L.w) + (size_t)c * bytes, bytes,
                                  smem_u32(w_full + slot));
                }
            }
        }
        __syncwarp();
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == F_EPI_WARPS) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, 512u);
    }
}

template <int CH>
cudaError_t launch_ch(const mrf::Params &p, int total_windows, cudaStream_t st)
{
    using G = mrf::Geo<CH>;
    uint32_t slot = 0;
    for (int l = 0; l < p.nlayers; ++l) slot = max(slot, mrf::chunk_bytes(p.L[l].k, G::S, CH));
    const size_t fixed = F_HEADER + f_tables_bytes<CH>() + 2 * (size_t)G::BUF;
    const size_t budget = 227 * 1024;
    int nslots = (int)((budget - fixed) / slot);
    if (nslots < 1) return cudaErrorInvalidConfiguration;
    if (nslots > F_MAX_SLOTS) nslots = F_MAX_SLOTS;
    const size_t smem = fixed + (size_t)nslots * slot;
    mrf_fused_kernel<CH><<<total_windows, F_THREADS, smem, st>>>(p, nslots, slot);
    return cudaGetLastError();
}

}  // namespace

cudaError_t mrf_fused_init()
{
    cudaError_t e;
    const int kMax = 227 * 1024;
    if ((e = cudaFuncSetAttribute(mrf_fused_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMax)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(mrf_fused_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMax)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(mrf_fused_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMax)) != cudaSuccess) return e;
    return cudaSuccess;
}

cudaError_t mrf_fused_launch(int CH, const mrf::Params &p, int total_windows, cudaStream_t st)
{
    switch (CH) {
        case 32:  return launch_ch<32>(p, total_windows, st);
        case 64:  return launch_ch<64>(p, total_windows, st);
        case 128: return launch_ch<128>(p, total_windows, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace zvx
