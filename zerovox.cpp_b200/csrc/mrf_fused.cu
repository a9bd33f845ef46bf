// mrf_fused.cu -- one HiFi-GAN MRF residual block (or a chain of its conv pairs) per launch,
// kept on chip for a whole time window.  Replaces HiFiGANResidualBlock
// (/root/reference/src/hifigan.cpp:74-185): 2 x (#pairs) ggml_conv_1d with leaky-ReLU, bias and
// the residual add in between, plus the branch sum / average of hifigan.cpp:300-315 in the
// final epilogue.  Geometry and the reasons for the swapped (weights = A operand) orientation
// are in mrf_fused.cuh.
//
// One CTA owns one window of WP time steps of one utterance (NCOL = 256 columns: one CTA per SM,
// 16 epilogue warps; NCOL = 128: two CTAs per SM, 8 epilogue warps each, so that one CTA's
// epilogue overlaps the other's MMAs):
//   epilogue warps  prologue + epilogues.  Warp w works on TMEM lane quarter w & 3 and on the 64
//              columns [64 (w >> 2), +64).
//   MMA warp   one elected lane issues, per layer, (k + S - 1) * CH/16 tcgen05.mma of shape
//              M=128 (weights window) x N=NCOL (positions) x K=16.
//   loader warp  cp.async.bulk of per-(layer, K-step) weight chunks into a ring.
// Tensor memory: columns [0,NCOL) = H accumulator (conv1 output), [NCOL,2 NCOL) = y.  The
// residual stream y stays in tensor memory in fp32 for the whole chain: conv2's MMAs accumulate
// directly on top of it, its bias is added when y is read (cumulative bias, host-prepared).
// Shared memory: two activation buffers (fp16, layouts per mrf_fused.cuh) that alternate as
// MMA B operand / epilogue destination, the staged scatter tables and the weight ring.
//
// Epilogue data path.  Interior windows (no utterance edge inside): tcgen05.ld.16x256b hands
// every thread the accumulator in the mma-fragment layout (row = lane/4 (+8), two adjacent
// columns per register pair), so after bias + leaky-ReLU + cvt.f16x2 a stmatrix.x4.trans stores
// four 8(channels) x 8(positions) blocks as 16-byte rows [position][8 channels] -- exactly the
// K-major operand rows the next MMA reads; row addresses come from the scatter table.  Windows
// that touch an utterance edge use a scalar path (tcgen05.ld.32x32b, one 2-byte store per
// element) that also applies the per-layer zero masking (SURVEY.md H-d).
#include <cstdio>

#include "mrf_fused.cuh"
#include "ptx_sm100.cuh"
#include "zvx_common.cuh"
#include "zvx_internal.h"

namespace zvx {

namespace {

constexpr int F_HEADER    = 512;
constexpr int F_MAX_SLOTS = 8;

__device__ __forceinline__ float lrelu_max(float x, float a)
{
    // max(x, a x) == ggml's max(x,0) + a min(x,0) for 0 < a < 1 (up to the sign of zero)
    return fmaxf(x, __fmul_rn(a, x));
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi)
{
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ void stmatrix_x4_trans(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d)
                 : "memory");
}
// 16 lanes x 64 columns; thread T gets for column group cg (8 columns): r[4cg+0/1] = (lane T/4,
// columns 8cg + 2(T%4), +1), r[4cg+2/3] = (lane T/4 + 8, same columns)
__device__ __forceinline__ void tmem_ld_16x256b_x8_nowait(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_16x256b_x8(uint32_t taddr, uint32_t (&r)[32])
{
    tmem_ld_16x256b_x8_nowait(taddr, r);
    tmem_wait_ld();
}
// 16 lanes x 16 columns (column groups cg = 0..1)
__device__ __forceinline__ void tmem_st_16x256b_x2(uint32_t taddr, const uint32_t (&r)[8])
{
    asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}

template <int CH, int NCOL>
struct FCfg {
    using G = mrf::Geo<CH, NCOL>;
    static constexpr int EPI_WARPS = NCOL / 16;            // 64 columns per warp, 4 lane quarters
    static constexpr int EPI       = EPI_WARPS * 32;
    // + one warpgroup: MMA warp, loader warp, two idle warps.  Registers are handed out per CTA in units of four
    // warps, so the two idle warps cost nothing, and a complete warpgroup can give its registers away
    // (setmaxnreg): the kernel is compiled for 96 (NCOL = 256) / 80 (NCOL = 128) registers per thread, the last
    // warpgroup shrinks to REGS_AUX and the epilogue warpgroups grow to REGS_EPI --
    // 640 x 96 = 512 x 112 + 128 x 32,  384 x 80 = 256 x 104 + 128 x 32.
    static constexpr int THREADS   = EPI + 128;
    static constexpr int CTAS      = NCOL == 128 ? 2 : 1;
    static constexpr int REGS_EPI  = NCOL == 128 ? 104 : 112;
    static constexpr int REGS_AUX  = 32;
    static constexpr int TBL_WORDS = G::S * NCOL;
    // utterance tables (seg_start, win_start: 1 + B entries each) are copied to shared memory when the
    // batch is small enough; every window looks its utterance up in them
    static constexpr int SEG_SMEM_MAX = 1024;
    __host__ __device__ static constexpr uint32_t seg_bytes(int B) { return B + 1 <= SEG_SMEM_MAX ? (uint32_t)((2 * (B + 1) * 4 + 511) & ~511) : 0u; }
    // compact (16-bit row unit) copies of the scatter tables of all layers stay resident: tbl0 + one per
    // non-final layer
    __host__ __device__ static constexpr uint32_t tbl_bytes(int nlayers) { return (uint32_t)((nlayers * TBL_WORDS * 2 + 511) & ~511); }
    static constexpr size_t   SMEM_BUDGET = NCOL == 128 ? 112 * 1024 : 227 * 1024;
};

// Windows that touch an utterance edge: after a buffer has been written, the rows of the time steps outside the
// utterance are overwritten with zeros (the convolution's zero padding).  buf: the buffer just written, d: the dilation
// its layout belongs to.  Called by all epilogue threads; not inlined (three call sites between hot code).
template <int CH, int NCOL>
__device__ __noinline__ void zero_outside_rows(uint8_t *buf, int d, int tw, int T, int tid)
{
    using C = FCfg<CH, NCOL>;
    using G = typename C::G;
    const int lo = max(0, -tw), hi = min(G::WP, T - tw);         // time steps [lo, hi) of the window are inside
    const int ninv = lo + (G::WP - hi);
    asm volatile("bar.sync 1, %0;" ::"n"(C::EPI) : "memory");          // every row has been written by its owner
    for (int i = tid; i < ninv * G::GROUPS; i += C::EPI) {
        const int g = i % G::GROUPS;
        int tau = i / G::GROUPS;
        if (tau >= lo) tau += hi - lo;
        const int unit = mrf::dest_unit(tau, d, G::WP, G::S, G::GROUPS, NCOL);
        *reinterpret_cast<uint4 *>(buf + (size_t)g * G::LBO_B + (size_t)unit * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
}

// FULL = false is the product instantiation: only the vector data path and the plain fp32 output (what every default
// launch uses).  FULL = true additionally carries the scalar reference path (flags bit 0, kept for the tests) and the
// general output phase (running branch sum / scale / fp16 hand-off).  They are separate kernels because the cold paths
// sit BETWEEN the hot ones in the instruction stream: the full kernel is 142 KB of SASS, the hot loop of the lean one
// fits the 32 KB instruction cache.
template <int CH, int NCOL, bool FULL>
__global__ void __launch_bounds__(FCfg<CH, NCOL>::THREADS, FCfg<CH, NCOL>::CTAS)
    mrf_fused_kernel(const mrf::Params p, const int nslots, const uint32_t slot_bytes)
{
    // compile-time false in the lean kernel
    const bool scalar_path = FULL && (p.flags & 1);
    using C = FCfg<CH, NCOL>;
    using G = typename C::G;
    constexpr int S = G::S;
    constexpr int TBL_WORDS = C::TBL_WORDS;
    constexpr int EPI_WARPS = C::EPI_WARPS;
    constexpr uint32_t LBO_B = G::LBO_B;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t *bars      = reinterpret_cast<uint64_t *>(smem);
    uint64_t *w_full    = bars;                        // [F_MAX_SLOTS]
    uint64_t *w_empty   = bars + F_MAX_SLOTS;          // [F_MAX_SLOTS]
    uint64_t *acc_full  = bars + 2 * F_MAX_SLOTS;
    uint64_t *act_ready = bars + 2 * F_MAX_SLOTS + 1;   // a layer's whole input is in shared memory
    uint64_t *act_half  = bars + 2 * F_MAX_SLOTS + 2;   // ... its even 16-channel K-steps are (see the MMA issuer)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + 256);
#ifdef ZVX_FUSED_PHASES
    // hand-off chain timestamps of CTA 0 (clock64 is per SM: comparable between warps): [0] commit of a layer issued,
    // [1] last epilogue warp published the first half, [2] ... the second half, [3] epilogue warp 0 woke on acc_full
    unsigned long long *ts = reinterpret_cast<unsigned long long *>(smem + 320);
    if (threadIdx.x < 4) ts[threadIdx.x] = 0ull;
#endif
    const uint32_t OFF_TBL = F_HEADER + C::seg_bytes(p.B);
    const int *seg_s    = reinterpret_cast<const int *>(smem + F_HEADER);      // [B + 1] seg_start, [B + 1] win_start
    const bool seg_in_smem = C::seg_bytes(p.B) != 0;
    uint16_t *tbl_s     = reinterpret_cast<uint16_t *>(smem + OFF_TBL);        // [nlayers][S][NCOL] row units
    const uint32_t OFF_BUF0 = OFF_TBL + C::tbl_bytes(p.nlayers);
    const uint32_t OFF_BUF1 = OFF_BUF0 + G::BUF;
    const uint32_t OFF_RING = OFF_BUF1 + G::BUF;

    const int tid  = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;

    const uint32_t smem_base = smem_u32(smem);
    const uint32_t buf0      = smem_base + OFF_BUF0;
    const uint32_t buf1      = smem_base + OFF_BUF1;
    const uint32_t ring      = smem_base + OFF_RING;

    // ---- persistent CTA: windows blockIdx.x, blockIdx.x + gridDim.x, ... of the launch ----
    struct Win { int T; size_t row0; int tw; bool interior; };
    auto window = [&](int win) {
        Win w;
        int wi, f0, f1;
        if (seg_in_smem) {
            const int *ws = seg_s + p.B + 1;
            int lo = 0, hi = p.B - 1;
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (ws[mid] <= win) lo = mid; else hi = mid - 1;
            }
            wi = win - ws[lo];
            f0 = seg_s[lo];
            f1 = seg_s[lo + 1];
        } else {
            const int u = find_segment_warp(p.win_start, p.B, win);
            wi = win - __ldg(p.win_start + u);
            f0 = __ldg(p.seg_start + u);
            f1 = __ldg(p.seg_start + u + 1);
        }
        w.T    = (f1 - f0) * p.rate;
        w.row0 = (size_t)f0 * p.rate;
        w.tw   = wi * p.valid - p.halo;                // time of window position 0 (may be < 0)
        w.interior = w.tw >= 0 && w.tw + G::WP <= w.T;
        return w;
    };
    const int nwin = p.total_windows;
    const int nl   = p.nlayers;

    // ---- one-time setup.  Every data row of a buffer that holds a position inside the window is
    //      rewritten by each layer before it is read; what must read as finite zeros are the guard
    //      rows and the data rows of positions beyond the window (never written) ----
    {
        constexpr int ROW_LO = G::WP / S;                      // first data row not fully covered by the window
        constexpr int NZ     = mrf::GUARD + (G::NROWS - mrf::GUARD - ROW_LO);   // rows to clear per (buffer, sub-buffer, group)
        for (int i = tid; i < 2 * S * G::GROUPS * NZ; i += C::THREADS) {
            const int sg = i / NZ, z = i % NZ;
            const int row = z < mrf::GUARD ? z : mrf::GUARD + ROW_LO + (z - mrf::GUARD);
            *reinterpret_cast<uint4 *>(smem + OFF_BUF0 + (size_t)sg * LBO_B + (size_t)row * 16) = make_uint4(0u, 0u, 0u, 0u);
        }
    }
    if (seg_in_smem) {
        int *dst = reinterpret_cast<int *>(smem + F_HEADER);
        for (int i = tid; i <= p.B; i += C::THREADS) {
            dst[i] = __ldg(p.seg_start + i);
            dst[p.B + 1 + i] = __ldg(p.win_start + i);
        }
    }
    for (int t = 0; t < nl; ++t) {
        const uint32_t *src = t == 0 ? p.tbl0 : p.L[t - 1].tbl;
        for (int i = tid; i < TBL_WORDS; i += C::THREADS) tbl_s[t * TBL_WORDS + i] = (uint16_t)(mrf::tbl_byte(__ldg(src + i)) >> 4);
    }
    if (tid == 0) {
        for (int s = 0; s < nslots; ++s) {
            mbar_init(smem_u32(w_full + s), 1);
            mbar_init(smem_u32(w_empty + s), 1);
        }
        mbar_init(smem_u32(acc_full), 1);
        mbar_init(smem_u32(act_ready), C::EPI);
        mbar_init(smem_u32(act_half), C::EPI);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == EPI_WARPS) tmem_alloc(smem_u32(tmem_slot), 2u * NCOL);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    // The two halves of tensor memory swap roles every window: y of window i lives where the conv1
    // accumulator of window i-1 lived, so that the next window's y can be written while the current
    // window's last conv still accumulates into its own y.
    auto ycol = [&](int iter) { return (uint32_t)((iter & 1) ? 0 : NCOL); };
    auto hcol = [&](int iter) { return (uint32_t)((iter & 1) ? NCOL : 0); };

    if (warp < EPI_WARPS) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(C::REGS_EPI));
        // =================== prologue + epilogues ===================
        const int quarter = warp & 3;
        const int colw    = (warp >> 2) * 64;          // this warp's 64 columns
        // ---- scalar-path coordinates: thread = accumulator row m ----
        const int m       = quarter * 32 + lane;
        const int s       = m / CH;
        const int oc      = m % CH;                    // row inside the CH block (kernel row order)
        const int gc      = mrf::row_to_chan(oc);      // the global channel it holds
        const uint32_t toff = (uint32_t)(oc >> 3) * LBO_B + (uint32_t)(oc & 7) * 2u;
        const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16);
        // ---- fragment-path coordinates: per 16-lane half lh, rows rb + lane/4 and rb + lane/4 + 8;
        //      over both halves a thread owns rows g, g+8, g+16, g+24 of its lane quarter = the four
        //      adjacent global channels c4 .. c4+3 (mrf::row_to_chan) of shift sQ ----
        const int mi = lane >> 3, r8 = lane & 7;       // stmatrix: this thread addresses row r8 of matrix mi
        const int sQ = (quarter * 32) / CH;
        const int c4 = (quarter * 32) % CH + 4 * (lane >> 2);

#ifdef ZVX_FUSED_PHASES
        const bool dbg = (p.flags & 2) && blockIdx.x == 0 && warp == 0;
        long long c_ld = 0, c_st = 0, c_l0 = 0;
#else
        constexpr bool dbg = false;
#endif
        // y window -> tensor memory columns [ybase, ybase + NCOL) (fp32), lrelu(y) -> buffer 0 (fp16).
        // Uses table slot 1; does NOT arrive on act_ready.
        // Interior windows, per 16-column quarter qc of this warp's 64 columns: 4 float4 loads per thread (its 4
        // channels x 4 columns); a warp instruction covers 4 time steps x one 128-byte line.  The loads are software-
        // pipelined: quarter 0 is requested by the caller one layer ahead (before the wait for the second-to-last
        // conv, where these warps idle anyway), quarter q + 1 before quarter q is placed, so that no L2 round trip
        // is exposed.
        // Addressing: tau = tauT + (compile-time offset); inside the window and the utterance iff tau - lo < span (one
        // unsigned compare), address = one per-window pointer + constant -- the loads are inlined five times and the
        // hot loop has to fit the instruction cache.
        const int tauT = S * (colw + 2 * (lane & 3)) + sQ;
        auto y_loads = [&](const Win &w, int qc, float4 (&f)[4]) {
            const int lo = max(0, -w.tw);
            const unsigned span = (unsigned)(min(G::WP, w.T - w.tw) - lo);
            const float *yq = p.y_in + ((ptrdiff_t)w.row0 + w.tw + tauT) * CH + c4;
            const int tl = tauT - lo;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int dt = S * (16 * qc + 8 * (i >> 1) + (i & 1));
                f[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if ((unsigned)(tl + dt) < span) f[i] = __ldg(reinterpret_cast<const float4 *>(yq + dt * CH));
            }
        };
        // Windows that touch an utterance edge take the same data path; afterwards the rows of the time steps outside
        // the utterance are overwritten with zeros (the convolution's zero padding, SURVEY.md H-d): typically a few
        // hundred 16-byte stores per CTA instead of a scalar pass over the whole window.  buf_off: the buffer just
        // written, d: the dilation its layout belongs to.
        auto zero_outside = [&](uint32_t buf_off, int d, int tw_, int T_) { zero_outside_rows<CH, NCOL>(smem + buf_off, d, tw_, T_, tid); };
        auto y_place = [&](int qc, const float4 (&f)[4], uint32_t ybase) {
            const int col0 = colw + 16 * qc;
            const uint16_t *tb = tbl_s + sQ * NCOL + col0;
#pragma unroll
            for (int lh = 0; lh < 2; ++lh) {
                const int rb = quarter * 32 + lh * 16;
                uint32_t v[8];
#pragma unroll
                for (int cg = 0; cg < 2; ++cg) {
                    v[4 * cg + 0] = __float_as_uint(lh ? f[2 * cg].z : f[2 * cg].x);
                    v[4 * cg + 1] = __float_as_uint(lh ? f[2 * cg + 1].z : f[2 * cg + 1].x);
                    v[4 * cg + 2] = __float_as_uint(lh ? f[2 * cg].w : f[2 * cg].y);
                    v[4 * cg + 3] = __float_as_uint(lh ? f[2 * cg + 1].w : f[2 * cg + 1].y);
                }
                tmem_st_16x256b_x2(tmem_base + ((uint32_t)rb << 16) + ybase + (uint32_t)col0, v);
                const uint32_t gbase = buf0 + (uint32_t)((rb % CH) >> 3) * LBO_B + (uint32_t)(mi & 1) * LBO_B;
                uint32_t h[4];
#pragma unroll
                for (int q = 0; q < 4; ++q)      // (cg = q / 2, row A/B = q & 1)
                    h[q] = pack_h2(lrelu_max(__uint_as_float(v[2 * q]), p.in_slope), lrelu_max(__uint_as_float(v[2 * q + 1]), p.in_slope));
                const uint32_t e = tb[8 * (mi >> 1) + r8];
                stmatrix_x4_trans(gbase + e * 16u, h[0], h[1], h[2], h[3]);
            }
        };
        auto prologue = [&](const Win &w, uint32_t ybase, bool preloaded, float4 (&f0)[4]) {
            if (!scalar_path) {
#ifdef ZVX_FUSED_PHASES
                if (dbg) c_l0 = clock64();
#endif
                if (!preloaded) y_loads(w, 0, f0);
                float4 f1[4];
                y_loads(w, 1, f1);
#ifdef ZVX_FUSED_PHASES
                if (dbg) {
                    asm volatile("" ::"f"(f0[0].x), "f"(f0[1].x), "f"(f0[2].x), "f"(f0[3].x) : "memory");
                    const long long c = clock64();
                    c_ld += c - c_l0;
                    c_l0 = c;
                }
#endif
                y_place(0, f0, ybase);
                y_loads(w, 2, f0);
                y_place(1, f1, ybase);
                y_loads(w, 3, f1);
                y_place(2, f0, ybase);
                y_place(3, f1, ybase);
                if (!w.interior) zero_outside(OFF_BUF0, p.L[0].d, w.tw, w.T);
            } else {
                const float *yin = p.y_in + w.row0 * CH + gc;
                const uint32_t *tb = p.tbl0 + s * NCOL;             // full entries (tau for the edge mask) from global
                uint8_t *dst = smem + OFF_BUF0 + toff;
#pragma unroll 1
                for (int b = 0; b < 2; ++b) {
                    const int col0 = colw + b * 32;
                    uint32_t v[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int tau = S * (col0 + i) + s;
                        const int t   = w.tw + tau;
                        float y = 0.f;
                        if (tau < G::WP && t >= 0 && t < w.T) y = __ldg(yin + (size_t)t * CH);
                        v[i] = __float_as_uint(y);
                    }
                    tmem_st32(tlane + ybase + (uint32_t)col0, v);
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const uint32_t e = __ldg(tb + col0 + i);
                        const __half h = __float2half_rn(lrelu_max(__uint_as_float(v[i]), p.in_slope));
                        if (e & mrf::TBL_VALID) *reinterpret_cast<__half *>(dst + mrf::tbl_byte(e)) = h;
                    }
                }
            }
#ifdef ZVX_FUSED_PHASES
            if (dbg) c_l0 = clock64();
#endif
            tmem_wait_st();
            fence_proxy_async_smem();
#ifdef ZVX_FUSED_PHASES
            if (dbg) c_st += clock64() - c_l0;
#endif
        };
        auto publish_half = [&]() {
            tc_fence_before_sync();
            mbar_arrive(smem_u32(act_half));
        };
        auto publish = [&]() {       // both barriers: everything this thread had to write is written
            tc_fence_before_sync();
            mbar_arrive(smem_u32(act_half));
            mbar_arrive(smem_u32(act_ready));
        };
        auto publish_rest = [&]() {  // after publish_half
            tc_fence_before_sync();
            mbar_arrive(smem_u32(act_ready));
        };

        // -DZVX_FUSED_PHASES + flags bit 1: cycle counts per phase of CTA 0 / warp 0, printed at exit
        long long c_pro = 0, c_wait = 0, c_drain = 0, c_final = 0, c_a = 0;
        [[maybe_unused]] long long c_top = 0, c_layers = 0, c_d3 = 0;
        [[maybe_unused]] const long long c_t0 = dbg ? clock64() : 0;
        int iter = 0;
        int win = blockIdx.x;
        Win wnext = {0, 0, 0, false};
        float4 ypre[4];                 // the next window's first 16-column quarter of y, requested one layer ahead
        bool have_pre = false;
        const bool preload = !(p.flags & 4);
        if (win < nwin) {
            wnext = window(win);
            prologue(wnext, ycol(0), false, ypre);
            publish();
        }
#ifdef ZVX_FUSED_PHASES
        const long long c_first = dbg ? clock64() - c_t0 : 0, c_ld_first = c_ld;
        // flags bit 3: every CTA reports its run time and how many of its windows touched an utterance edge
        const bool dbg_cta = (p.flags & 8) && warp == 0 && lane == 0;
        const long long c_cta0 = dbg_cta ? clock64() : 0;
        int n_edge = 0;
#endif
#pragma unroll 1
        for (; win < nwin; win += gridDim.x, ++iter) {
            const Win w = wnext;
            const int tw = w.tw, T = w.T;
            const bool interior = w.interior;
            const bool has_next = win + (int)gridDim.x < nwin;
#ifdef ZVX_FUSED_PHASES
            n_edge += interior ? 0 : 1;
#endif
#pragma unroll 1
            for (int l = 0; l < nl; ++l) {
                const mrf::Layer &L = p.L[l];
                const bool last = l == nl - 1;
                const uint32_t acc_col = L.accumulate ? ycol(iter) : hcol(iter);
                const uint32_t gl = (uint32_t)(iter * nl + l);      // completions of acc_full before this one
#ifdef ZVX_FUSED_PHASES
                const long long c_it0 = dbg ? clock64() : 0;
#endif
                if (has_next && l == nl - 2) {
                    // the next window's coordinates, and the first half of its y on the way while these warps
                    // wait for this conv and drain it
                    wnext = window(win + (int)gridDim.x);
                    have_pre = preload && !scalar_path;
                    if (have_pre) y_loads(wnext, 0, ypre);
                }
#ifdef ZVX_FUSED_PHASES
                if (dbg) c_top += clock64() - c_it0;
#endif
                if (last && has_next) {
                    // While the last conv of this window accumulates into y, bring in the NEXT window: its
                    // y goes to the (now idle) conv1 accumulator columns, lrelu(y) to buffer 0 (the last
                    // layer reads buffer 1; nlayers is even).  Published after this window's y is read.
                    if (dbg) c_a = clock64();
                    prologue(wnext, hcol(iter), have_pre, ypre);
                    if (dbg) c_pro += clock64() - c_a;
                }
                // this thread's four biases of the layer (rows g, g+8 of both 16-lane halves of its lane quarter),
                // requested before the wait: their L2 round trip used to sit between the wait and the first add
                const int rq = (quarter * 32) % CH + (lane >> 2);
                const float4 b4 = make_float4(__ldg(L.bias + rq), __ldg(L.bias + rq + 8), __ldg(L.bias + rq + 16), __ldg(L.bias + rq + 24));
                if (dbg) c_a = clock64();
                mbar_wait(smem_u32(acc_full), gl & 1u, p.err_flag);
                tc_fence_after_sync();
                if (dbg) { const long long c_b = clock64(); c_wait += c_b - c_a; c_a = c_b; }
#ifdef ZVX_FUSED_PHASES
                if (dbg && lane == 0) { ts[3] = (unsigned long long)c_a; c_d3 += c_a - (long long)ts[0]; }
#endif
                if (!last) {
                    const uint32_t obuf_off = (l & 1) ? OFF_BUF0 : OFF_BUF1;
                    const float slope = L.out_slope;
                    if (!scalar_path) {
#pragma unroll 1
                        for (int lh = 0; lh < 2; ++lh) {
                            const int rb  = quarter * 32 + lh * 16;
                            const int sA  = rb / CH;
                            const float bA = lh ? b4.z : b4.x, bB = lh ? b4.w : b4.y;
                            const uint16_t *tb = tbl_s + (l + 1) * TBL_WORDS + sA * NCOL + colw;
                            uint32_t r[32];
                            tmem_ld_16x256b_x8(tmem_base + ((uint32_t)rb << 16) + acc_col + (uint32_t)colw, r);
                            const uint32_t gbase = smem_base + obuf_off + (uint32_t)((rb % CH) >> 3) * LBO_B + (uint32_t)(mi & 1) * LBO_B;
#pragma unroll
                            for (int pr = 0; pr < 4; ++pr) {
                                uint32_t f[4];
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    const int i0 = 8 * pr + 2 * q;
                                    const float b = (q & 1) ? bB : bA;
                                    f[q] = pack_h2(lrelu_max(__fadd_rn(__uint_as_float(r[i0]), b), slope),
                                                   lrelu_max(__fadd_rn(__uint_as_float(r[i0 + 1]), b), slope));
                                }
                                const uint32_t e = tb[8 * (2 * pr + (mi >> 1)) + r8];
                                stmatrix_x4_trans(gbase + e * 16u, f[0], f[1], f[2], f[3]);
                            }
                            if (lh == 0 && interior) {
                                // rows [rb, rb + 16) of every lane quarter = the even K-steps of the next layer's
                                // operand: its MMAs over those start while the odd half is still being drained
                                fence_proxy_async_smem();
#ifdef ZVX_FUSED_PHASES
                                if ((p.flags & 2) && blockIdx.x == 0 && lane == 0) atomicMax(&ts[1], (unsigned long long)clock64());
#endif
                                publish_half();
                            }
                        }
                        if (interior) {
                            fence_proxy_async_smem();
#ifdef ZVX_FUSED_PHASES
                            if ((p.flags & 2) && blockIdx.x == 0 && lane == 0) atomicMax(&ts[2], (unsigned long long)clock64());
#endif
                            publish_rest();
                        } else {
                            zero_outside(obuf_off, p.L[l + 1].d, tw, T);
                            fence_proxy_async_smem();
                            publish();
                        }
                    } else {
                        const float bias = __ldg(L.bias + oc);
                        const uint32_t *tb = L.tbl + s * NCOL;              // full entries from global
                        uint8_t *dst = smem + obuf_off + toff;
#pragma unroll 1
                        for (int b = 0; b < 2; ++b) {
                            const int col0 = colw + b * 32;
                            uint32_t r[32];
                            tmem_ld32(tlane + acc_col + (uint32_t)col0, r);
#pragma unroll
                            for (int i = 0; i < 32; ++i) {
                                const uint32_t e = __ldg(tb + col0 + i);
                                float v = lrelu_max(__fadd_rn(__uint_as_float(r[i]), bias), slope);
                                const int t = tw + mrf::tbl_tau(e);
                                if (t < 0 || t >= T) v = 0.f;
                                if (e & mrf::TBL_VALID) *reinterpret_cast<__half *>(dst + mrf::tbl_byte(e)) = __float2half_rn(v);
                            }
                        }
                        fence_proxy_async_smem();
                        publish();
                    }
                    if (dbg) c_drain += clock64() - c_a;
                } else {
                    // final: y (+ running branch sum) (* 1/num_blocks) -> global fp32.  Both 32-column
                    // batches are pulled out of tensor memory first, then the next window is published
                    // (its first conv may overwrite these columns), then the global traffic follows.
                    if (!scalar_path) {
                        // fragment layout again: this thread's 4 adjacent channels x 16 columns leave as float4
                        uint32_t r0[32], r1[32];
                        tmem_ld_16x256b_x8_nowait(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc_col + (uint32_t)colw, r0);
                        tmem_ld_16x256b_x8_nowait(tmem_base + ((uint32_t)(quarter * 32 + 16) << 16) + acc_col + (uint32_t)colw, r1);
                        tmem_wait_ld();
                        if (has_next) publish();
                        // time of this thread's column i: tau0 + S * (8 (i / 2) + (i & 1)); inside the window's valid range and
                        // the utterance iff tau - lo < span (one unsigned compare, addresses = one pointer + constants)
                        const int tau0 = S * (colw + 2 * (lane & 3)) + sQ;
                        const int lo = p.halo;
                        const unsigned span = (unsigned)(min(p.halo + p.valid, T - tw) - lo);
                        const ptrdiff_t base = ((ptrdiff_t)w.row0 + tw + tau0) * CH + c4;
                        if (!FULL || (p.out && !p.out16 && !p.acc_in && !p.has_scale)) {
                            // the common case: y (+ bias) leaves as fp32, nothing else
                            float *oq = p.out + base;
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const int dt = S * (8 * (i >> 1) + (i & 1));
                                if ((unsigned)(tau0 + dt - lo) < span) {
                                    const int i0 = 4 * (i >> 1) + (i & 1);
                                    *reinterpret_cast<float4 *>(oq + dt * CH) =
                                        make_float4(__fadd_rn(__uint_as_float(r0[i0]), b4.x), __fadd_rn(__uint_as_float(r0[i0 + 2]), b4.y),
                                                    __fadd_rn(__uint_as_float(r1[i0]), b4.z), __fadd_rn(__uint_as_float(r1[i0 + 2]), b4.w));
                                }
                            }
                        } else {
                        float *oq = p.out ? p.out + base : nullptr;
                        uint16_t *hq = p.out16 ? p.out16 + base : nullptr;
                        const float *aq = p.acc_in ? p.acc_in + base : nullptr;
                        const float *aq2 = p.acc_in2 ? p.acc_in2 + base : nullptr;
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int dt = S * (8 * (i >> 1) + (i & 1));
                            if ((unsigned)(tau0 + dt - lo) < span) {
                                const int i0 = 4 * (i >> 1) + (i & 1);
                                float4 v = make_float4(__fadd_rn(__uint_as_float(r0[i0]), b4.x), __fadd_rn(__uint_as_float(r0[i0 + 2]), b4.y),
                                                       __fadd_rn(__uint_as_float(r1[i0]), b4.z), __fadd_rn(__uint_as_float(r1[i0 + 2]), b4.w));
                                if (aq) {
                                    float4 a = *reinterpret_cast<const float4 *>(aq + dt * CH);
                                    if (aq2) {       // (y_0 + y_1) + y_2: the reference's order (hifigan.cpp:300-311)
                                        const float4 a2 = *reinterpret_cast<const float4 *>(aq2 + dt * CH);
                                        a = make_float4(__fadd_rn(a.x, a2.x), __fadd_rn(a.y, a2.y), __fadd_rn(a.z, a2.z), __fadd_rn(a.w, a2.w));
                                    }
                                    v = make_float4(__fadd_rn(a.x, v.x), __fadd_rn(a.y, v.y), __fadd_rn(a.z, v.z), __fadd_rn(a.w, v.w));
                                }
                                if (p.has_scale) v = make_float4(__fmul_rn(v.x, p.scale), __fmul_rn(v.y, p.scale), __fmul_rn(v.z, p.scale), __fmul_rn(v.w, p.scale));
                                if (oq) *reinterpret_cast<float4 *>(oq + dt * CH) = v;
                                if (hq) {
                                    uint2 h;
                                    h.x = pack_h2(lrelu_max(v.x, p.out16_slope), lrelu_max(v.y, p.out16_slope));
                                    h.y = pack_h2(lrelu_max(v.z, p.out16_slope), lrelu_max(v.w, p.out16_slope));
                                    *reinterpret_cast<uint2 *>(hq + dt * CH) = h;
                                }
                            }
                        }
                        }
                    } else {
                    const float bias = __ldg(L.bias + oc);
                    float *out = p.out ? p.out + w.row0 * CH + gc : nullptr;
                    uint16_t *out16 = p.out16 ? p.out16 + w.row0 * CH + gc : nullptr;
                    const float *ain = p.acc_in ? p.acc_in + w.row0 * CH + gc : nullptr;
                    const float *ain2 = p.acc_in2 ? p.acc_in2 + w.row0 * CH + gc : nullptr;
                    uint32_t r0[32], r1[32];
                    tmem_ld32(tlane + acc_col + (uint32_t)colw, r0);
                    tmem_ld32(tlane + acc_col + (uint32_t)(colw + 32), r1);
                    if (has_next) publish();
#pragma unroll
                    for (int b = 0; b < 2; ++b) {
                        const int col0 = colw + b * 32;
                        const uint32_t (&r)[32] = b ? r1 : r0;
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            float a[16];
                            if (ain) {
#pragma unroll
                                for (int i = 0; i < 16; ++i) {
                                    const int tau = S * (col0 + 16 * h + i) + s;
                                    const int t   = tw + tau;
                                    const bool ok = tau >= p.halo && tau < p.halo + p.valid && t < T;
                                    a[i] = ok ? ain[(size_t)t * CH] : 0.f;
                                    if (ok && ain2) a[i] = __fadd_rn(a[i], ain2[(size_t)t * CH]);
                                }
                            }
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const int tau = S * (col0 + 16 * h + i) + s;
                                const int t   = tw + tau;
                                if (tau >= p.halo && tau < p.halo + p.valid && t < T) {
                                    float v = __fadd_rn(__uint_as_float(r[16 * h + i]), bias);
                                    if (ain) v = __fadd_rn(a[i], v);
                                    if (p.has_scale) v = __fmul_rn(v, p.scale);
                                    if (out) out[(size_t)t * CH] = v;
                                    if (out16) out16[(size_t)t * CH] = __half_as_ushort(__float2half_rn(lrelu_max(v, p.out16_slope)));
                                }
                            }
                        }
                    }
                    }
                    if (dbg) c_final += clock64() - c_a;
                }
#ifdef ZVX_FUSED_PHASES
                if (dbg) c_layers += clock64() - c_it0;
#endif
            }
        }
#ifdef ZVX_FUSED_PHASES
        if (dbg_cta) printf("mrf_cta CH=%d NCOL=%d k=%d nl=%d cta=%d windows=%d edge=%d cycles=%lld\n", CH, NCOL, p.L[0].k, nl, (int)blockIdx.x, iter, n_edge, clock64() - c_cta0);
        if (dbg && lane == 0) printf("mrf_chain commit issued -> epilogue awake %lld (cycles, summed over all layers of CTA 0)\n", c_d3);
        if (dbg && lane == 0)
            printf("mrf_fused CH=%d NCOL=%d k=%d nl=%d windows=%d: total %lld  prologue %lld  wait_mma %lld  drain %lld  final %lld  | prologue: loads %lld  wait_st+fence %lld  first prologue (tensor pipe idle) %lld of which loads %lld | layer iterations %lld of which next-window lookup + preload issue %lld (cycles, CTA 0 warp 0)\n",
                   CH, NCOL, p.L[0].k, nl, iter, clock64() - c_t0, c_pro, c_wait, c_drain, c_final, c_ld, c_st, c_first, c_ld_first, c_layers, c_top);
#endif
    } else {
    // one instruction for the whole last warpgroup (.aligned), then the roles split
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(C::REGS_AUX));
    if (warp == EPI_WARPS) {
        // =================== MMA issuer ===================
        const uint32_t leader = elect_one();
        const uint32_t idesc  = make_idesc_mn(128, NCOL);
#ifdef ZVX_FUSED_PHASES
        const bool dbg = (p.flags & 2) && blockIdx.x == 0;
#else
        constexpr bool dbg = false;
#endif
        long long c_act = 0, c_w = 0, c_a = 0, c_t0 = dbg ? clock64() : 0;
        [[maybe_unused]] long long c_d1 = 0, c_d2 = 0, c_d4 = 0, c_d5 = 0, c_d6 = 0, c_d7 = 0, c_n = 0, t_half = 0, t_ready = 0;
        int it = 0, iter = 0;
#pragma unroll 1
        for (int win = blockIdx.x; win < nwin; win += gridDim.x, ++iter) {
#ifdef ZVX_FUSED_PHASES
            const bool interior_dbg = dbg && window(win).interior;
#endif
#pragma unroll 1
            for (int l = 0; l < nl; ++l) {
                const mrf::Layer &L = p.L[l];
                const int k  = L.k;
                const int nj = k + S - 1;
                const uint32_t lbo_a = (uint32_t)mrf::tap_blocks(k, S) * CH * 16u;
                const uint32_t ibuf  = (l & 1) ? buf1 : buf0;
                const uint32_t dcol  = tmem_base + (L.accumulate ? ycol(iter) : hcol(iter));
                // Step j = 0 of every K-step: the A window starts at tap block k + S - 2 and moves down one block
                // per step; the B operand starts at position -c = sub-buffer q0, row ro0 and moves up one position
                // per step (next sub-buffer; after the last one, the first sub-buffer one row further).  Only the
                // 14-bit start-address fields of the two descriptors change, by constants: no per-step index
                // arithmetic in the issuing thread (it was the limit: 140-215 cycles per MMA instead of 128).
                int q0, ro0;
                mrf::b_step(k, S, 0, q0, ro0);
                const uint32_t a_fix = ((lbo_a >> 4) & 0x3FFFu) << 16;
                const uint32_t b_fix = ((LBO_B >> 4) & 0x3FFFu) << 16;
                const uint32_t a_off0 = (uint32_t)(k + S - 2) * (CH * 16u);
                const uint32_t b_off0 = (uint32_t)q0 * G::SUB + (uint32_t)(ro0 * 16) + (uint32_t)mrf::GUARD * 16u;
                constexpr uint64_t DESC_HI = ((uint64_t)(128u >> 4) | ((uint64_t)1 << 14)) << 32;   // SBO = 128 B, version 1
                // K-steps in the order even, then odd: the epilogue warps publish the even ones (first 16-lane
                // half of every lane quarter) before they drain the odd ones
#pragma unroll 1
                for (int ci = 0; ci < G::KSTEPS; ++ci, ++it) {
                    const int c = ci < G::KSTEPS / 2 ? 2 * ci : 2 * (ci - G::KSTEPS / 2) + 1;
                    if (ci == 0 || ci == G::KSTEPS / 2) {
                        if (dbg) c_a = clock64();
                        mbar_wait(smem_u32(ci == 0 ? act_half : act_ready), (uint32_t)(iter * nl + l) & 1u, p.err_flag);
                        tc_fence_after_sync();
                        if (dbg) c_act += clock64() - c_a;
#ifdef ZVX_FUSED_PHASES
                        if (dbg && l > 0 && interior_dbg) {
                            const long long now = clock64();
                            if (ci == 0) {
                                c_d4 += (long long)ts[1] - (long long)ts[3];      // acc_full wake -> first half published
                                c_d5 += now - (long long)ts[1];                   // ... -> this warp awake
                                c_n += 1;
                                t_half = now;
                            } else {
                                c_d1 += now - t_half;                             // even K-steps issued + wait for the second half
                                c_d6 += (long long)ts[2] - (long long)ts[1];      // second half drained
                                c_d7 += now - (long long)ts[2];
                                t_ready = now;
                            }
                        }
#endif
                    }
                    const int slot = it % nslots;
                    const uint32_t ph = (uint32_t)(it / nslots) & 1u;
                    if (dbg) c_a = clock64();
                    mbar_wait(smem_u32(w_full + slot), ph, p.err_flag);
                    tc_fence_after_sync();
                    if (dbg) c_w += clock64() - c_a;
                    if (leader) {
                        uint32_t a_lo = (((ring + (uint32_t)slot * slot_bytes + a_off0) & 0x3FFFFu) >> 4) | a_fix;
                        uint32_t b_lo = (((ibuf + (uint32_t)(2 * c) * LBO_B + b_off0) & 0x3FFFFu) >> 4) | b_fix;
                        uint32_t acc = (L.accumulate || ci > 0) ? 1u : 0u;
                        int q = q0;
#pragma unroll 1
                        for (int j = 0; j < nj; ++j) {
                            umma_f16(dcol, DESC_HI | a_lo, DESC_HI | b_lo, idesc, acc);
                            acc = 1u;
                            a_lo -= (uint32_t)CH;                           // one tap block = CH rows of 16 bytes
                            b_lo += (uint32_t)(G::SUB >> 4);
                            if (++q == S) { q = 0; b_lo -= (uint32_t)(S * (G::SUB >> 4) - 1); }
                        }
                        umma_commit(smem_u32(w_empty + slot));
                    }
                    __syncwarp();
                }
                if (leader) umma_commit(smem_u32(acc_full));
#ifdef ZVX_FUSED_PHASES
                if (dbg) {
                    const long long now = clock64();
                    if (lane == 0) ts[0] = (unsigned long long)now;
                    if (l > 0 && interior_dbg) c_d2 += now - t_ready;            // odd K-steps issued
                }
#endif
                __syncwarp();
            }
        }
#ifdef ZVX_FUSED_PHASES
        if (dbg && lane == 0)
            printf("mrf_chain per hand-off (layers 1.., %lld samples): awake->half published %lld | ->MMA warp awake %lld | even K issue + wait %lld (second half drained after %lld, +%lld to wake) | odd K issue %lld\n",
                   c_n, c_d4 / max(c_n, 1LL), c_d5 / max(c_n, 1LL), c_d1 / max(c_n, 1LL), c_d6 / max(c_n, 1LL), c_d7 / max(c_n, 1LL), c_d2 / max(c_n, 1LL));
#endif
        if (dbg && lane == 0)
            printf("mrf_fused MMA warp: total %lld  wait_activations %lld  wait_weights %lld (cycles, CTA 0)\n", clock64() - c_t0, c_act, c_w);
    } else if (warp == EPI_WARPS + 1) {
        // =================== weight loader + L2 prefetcher ===================
        int it = 0;
#pragma unroll 1
        for (int win = blockIdx.x; win < nwin; win += gridDim.x) {
            // L2 prefetch of the inputs of this CTA's NEXT window (requested a whole window ahead of
            // the prologue that reads them): the y rows and, for the final phase, the running sum
            const int nw = win + (int)gridDim.x;
            if (p.prefetch && nw < nwin) {
                const Win wn = window(nw);
                const int lo = max(wn.tw, 0), hi = min(wn.tw + G::WP, wn.T);
                const char *ybase = reinterpret_cast<const char *>(p.y_in + (wn.row0 + (size_t)lo) * CH);
                if (p.prefetch == 2) {
                    // the window's rows are contiguous: one bulk prefetch request
                    if (lane == 0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ybase), "r"((uint32_t)((hi - lo) * (CH * 4))) : "memory");
                } else {
                    const int nlines = (hi - lo) * (CH * 4) / 128;
                    for (int i = lane; i < nlines; i += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(ybase + (size_t)i * 128));
                }
                for (int k = 0; k < 2; ++k) {
                    const float *acc = k ? p.acc_in2 : p.acc_in;
                    if (!acc) continue;
                    const int alo = max(wn.tw + p.halo, 0), ahi = min(wn.tw + p.halo + p.valid, wn.T);
                    const char *abase = reinterpret_cast<const char *>(acc + (wn.row0 + (size_t)alo) * CH);
                    const int alines = (ahi - alo) * (CH * 4) / 128;
                    for (int i = lane; i < alines; i += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(abase + (size_t)i * 128));
                }
            }
            if (lane == 0) {
                for (int l = 0; l < nl; ++l) {
                    const mrf::Layer &L = p.L[l];
                    const uint32_t bytes = mrf::chunk_bytes(L.k, S, CH);
                    for (int ci = 0; ci < G::KSTEPS; ++ci, ++it) {
                        const int c = ci < G::KSTEPS / 2 ? 2 * ci : 2 * (ci - G::KSTEPS / 2) + 1;   // the issuer's order
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
    }
    }   // last warpgroup (its other two warps idle: they only exist so that the warpgroup is complete)

    tc_fence_before_sync();
    __syncthreads();
    if (warp == EPI_WARPS) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, 2u * NCOL);
    }
}

template <int CH, int NCOL>
cudaError_t launch_cfg(const mrf::Params &p, int total_windows, cudaStream_t st)
{
    using C = FCfg<CH, NCOL>;
    uint32_t slot = 0;
    for (int l = 0; l < p.nlayers; ++l) slot = max(slot, mrf::chunk_bytes(p.L[l].k, C::G::S, CH));
    const size_t fixed = F_HEADER + C::seg_bytes(p.B) + C::tbl_bytes(p.nlayers) + 2 * (size_t)C::G::BUF;
    if (fixed + slot > C::SMEM_BUDGET) return cudaErrorInvalidConfiguration;
    int nslots = (int)((C::SMEM_BUDGET - fixed) / slot);
    if (nslots > F_MAX_SLOTS) nslots = F_MAX_SLOTS;
    const size_t smem = fixed + (size_t)nslots * slot;
    const int resident = p.resident_ctas > 0 ? p.resident_ctas : total_windows;
    const int grid = total_windows < resident ? total_windows : resident;
    mrf::Params q = p;
    q.total_windows = total_windows;
    const bool lean = !(p.flags & 1) && p.out && !p.out16 && !p.acc_in && !p.has_scale && !(p.flags & 16);
    if (lean) mrf_fused_kernel<CH, NCOL, false><<<grid, C::THREADS, smem, st>>>(q, nslots, slot);
    else      mrf_fused_kernel<CH, NCOL, true><<<grid, C::THREADS, smem, st>>>(q, nslots, slot);
    return cudaGetLastError();
}

template <int CH, int NCOL>
cudaError_t init_cfg()
{
    const cudaError_t e = cudaFuncSetAttribute(mrf_fused_kernel<CH, NCOL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FCfg<CH, NCOL>::SMEM_BUDGET);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(mrf_fused_kernel<CH, NCOL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FCfg<CH, NCOL>::SMEM_BUDGET);
}

}  // namespace

cudaError_t mrf_fused_init()
{
    cudaError_t e;
    if ((e = init_cfg<32, 256>()) != cudaSuccess) return e;
    if ((e = init_cfg<64, 256>()) != cudaSuccess) return e;
    if ((e = init_cfg<128, 256>()) != cudaSuccess) return e;
    if ((e = init_cfg<32, 128>()) != cudaSuccess) return e;
    if ((e = init_cfg<64, 128>()) != cudaSuccess) return e;
    return cudaSuccess;
}

bool mrf_fused_supported(int CH, int ncol)
{
    return (ncol == 256 && (CH == 32 || CH == 64 || CH == 128)) || (ncol == 128 && (CH == 32 || CH == 64));
}

cudaError_t mrf_fused_launch(int CH, const mrf::Params &p, int total_windows, cudaStream_t st)
{
    if (p.ncol == 256) {
        switch (CH) {
            case 32:  return launch_cfg<32, 256>(p, total_windows, st);
            case 64:  return launch_cfg<64, 256>(p, total_windows, st);
            case 128: return launch_cfg<128, 256>(p, total_windows, st);
        }
    } else if (p.ncol == 128) {
        switch (CH) {
            case 32:  return launch_cfg<32, 128>(p, total_windows, st);
            case 64:  return launch_cfg<64, 128>(p, total_windows, st);
        }
    }
    return cudaErrorInvalidValue;
}

}  // namespace zvx
